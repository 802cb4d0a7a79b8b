"""``pysparse`` import-path alias for pysparse_b200.

Lets unmodified PySparse user code (``from pysparse.sparse import spmatrix``,
``from pysparse.itsolvers.krylov import pcg``, ``from pysparse.precon import
precon``, ``from pysparse.tools import poisson``) run on the B200 implementation.
Only the SpMV / Krylov / preconditioner path is provided.
"""
import importlib
import sys

import pysparse_b200

__version__ = pysparse_b200.__version__

for _name in ("sparse", "sparse.spmatrix", "sparse.pysparseMatrix", "sparse.sparseMatrix", "itsolvers", "itsolvers.krylov", "itsolvers.itsolvers_util",
              "precon", "precon.precon", "eigen", "eigen.jdsym", "tools", "tools.poisson", "tools.poisson_vec", "tools.spmatrix_util",
              "tools.sptime", "tools.gallery"):
    _mod = importlib.import_module("pysparse_b200." + _name)
    sys.modules["pysparse." + _name] = _mod
    if "." not in _name:
        globals()[_name] = _mod
