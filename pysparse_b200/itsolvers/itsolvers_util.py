"""Object-oriented front end to the Krylov solvers.

Python-3 restatement of ``pysparse/itsolvers/itsolvers_util.py:13-152``: same class names,
``Solver(A, debug=False)`` construction (``A`` an ``ll_mat`` / ``csr_mat`` / ``sss_mat`` or a
``PysparseMatrix``, which is unwrapped), ``solve(b, x, tol, maxit, K=None)`` with a zero
initial guess, the counters ``nofCalled``, ``totalIterations``, ``lastIterations``,
``lastInfo``, and -- as in the reference -- a ``RuntimeError`` carrying ``(message, info,
iter, relres)`` when the solver reports ``info < 0``.  Additions that do not change that
protocol: a ``precon=`` keyword at construction (default ``K`` of ``solve``), the attributes
``itn``, ``resNorm``, ``converged``, and ``solve`` returns ``(info, iter, relres)``.
"""
from . import krylov

__all__ = ["ItSolver", "Pcg", "Minres", "Qmrs", "Cgs", "Bicgstab", "Gmres",
           "pcg", "minres", "qmrs", "cgs", "bicgstab", "gmres"]


class ItSolver:
    "Abstract base class for iterative solver classes."

    itsolver = None
    name = "Generic"

    def __init__(self, A, debug=False, **kwargs):
        from ..sparse.pysparseMatrix import PysparseMatrix
        self.A = A.matrix if isinstance(A, PysparseMatrix) else A
        self.matrix = self.A
        self.precon = kwargs.get("precon", None)
        self.debug = debug
        self.nofCalled = 0
        self.totalIterations = 0
        self.lastIterations = 0
        self.lastInfo = 0
        self.itn = 0
        self.resNorm = 0.0
        self.converged = False

    def set_name(self, name):
        self.name = name

    def set_solver(self, solver):
        self.itsolver = solver

    def solve(self, b, x, tol, maxit, K=None):
        "Solve Ax = b iteratively with zero initial guess"
        if not self.itsolver:
            raise NotImplementedError("This class cannot be used directly.")
        if K is None:
            K = self.precon
        x[:] = 0
        info, it, relres = self.itsolver(self.A, b, x, tol, maxit, K)
        self.nofCalled += 1
        self.totalIterations += it
        self.lastIterations = self.itn = it
        self.lastInfo = info
        self.resNorm = relres
        self.converged = info >= 0
        if self.debug:
            print("iterative solver returned:", info, it, relres)
        if info < 0:
            raise RuntimeError("iterative solver %s returned error code %d" % (self.__class__.__name__, info),
                               info, it, relres)
        return info, it, relres

    def __call__(self, *args, **kwargs):
        return self.solve(*args, **kwargs)

    def __str__(self):
        s = "<%s.%s instance>\n\n" % (self.__class__.__module__, self.__class__.__name__)
        for name in ("nofCalled", "totalIterations", "lastIterations", "lastInfo"):
            s += "   %s: %s\n" % (name, getattr(self, name))
        return s + "\n"

    __repr__ = __str__


class Pcg(ItSolver):
    name = "Pcg"
    itsolver = staticmethod(krylov.pcg)


class Minres(ItSolver):
    name = "Minres"
    itsolver = staticmethod(krylov.minres)


class Qmrs(ItSolver):
    name = "Qmrs"
    itsolver = staticmethod(krylov.qmrs)


class Cgs(ItSolver):
    name = "Cgs"
    itsolver = staticmethod(krylov.cgs)


class Bicgstab(ItSolver):
    name = "Bicgstab"
    itsolver = staticmethod(krylov.bicgstab)


class Gmres(ItSolver):
    name = "Gmres"
    itsolver = staticmethod(krylov.gmres)


# module-level functions of the reference (itsolvers_util.py:128-152; deprecated there in
# favour of the classes): plain pass-throughs to the krylov module
def pcg(*args, **kwargs):
    return krylov.pcg(*args, **kwargs)


def minres(*args, **kwargs):
    return krylov.minres(*args, **kwargs)


def qmrs(*args, **kwargs):
    return krylov.qmrs(*args, **kwargs)


def cgs(*args, **kwargs):
    return krylov.cgs(*args, **kwargs)


def bicgstab(*args, **kwargs):
    return krylov.bicgstab(*args, **kwargs)


def gmres(*args, **kwargs):
    return krylov.gmres(*args, **kwargs)
