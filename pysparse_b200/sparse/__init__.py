"""pysparse_b200.sparse -- drop-in for ``pysparse.sparse``: the ``spmatrix`` extension module
(``ll_mat`` / ``csr_mat`` / ``sss_mat``) and the ``PysparseMatrix`` wrapper classes."""
from . import spmatrix  # noqa: F401  (C extension; ImportError if the native build is missing)
from .pysparseMatrix import (PysparseIdentityMatrix, PysparseMatrix, PysparseMatrix4Scipy,  # noqa: F401
                             PysparseSpDiagsMatrix, SparseMatrix)
