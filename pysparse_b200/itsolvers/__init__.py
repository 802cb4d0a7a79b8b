"""pysparse_b200.itsolvers -- drop-in for ``pysparse.itsolvers`` (module ``krylov`` + OO wrappers)."""
from . import krylov  # noqa: F401
from .itsolvers_util import ItSolver, Pcg, Minres, Qmrs, Cgs, Bicgstab, Gmres  # noqa: F401
