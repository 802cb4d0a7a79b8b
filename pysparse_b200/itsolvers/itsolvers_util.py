"""Object-oriented front end to the Krylov solvers.

Python-3 restatement of ``pysparse/itsolvers/itsolvers_util.py:13-125``: same
class names, same ``solve(rhs, x, tol, maxit)`` protocol (x is zeroed first,
``info < 0`` is recorded, not raised unless ``debug``), same attributes
(``nofCalled``, ``totalIterations``, ``lastIterations``, ``lastInfo``,
``itn``, ``resNorm``, ``converged``).
"""
import numpy as np

from . import krylov

__all__ = ["ItSolver", "Pcg", "Minres", "Qmrs", "Cgs", "Bicgstab", "Gmres"]


class ItSolver:
    """Abstract base class of the iterative solvers."""

    itsolver = None
    name = "Generic"

    def __init__(self, matrix, **kwargs):
        self.matrix = matrix
        self.type = self.name
        self.precon = kwargs.get("precon", None)
        self.debug = kwargs.get("debug", False)
        self.nofCalled = 0
        self.totalIterations = 0
        self.lastIterations = 0
        self.lastInfo = 0
        self.itn = 0
        self.resNorm = 0.0
        self.converged = False

    def solve(self, rhs, x, tol, maxit, K=None):
        """Solve A x = rhs to relative tolerance ``tol`` in at most ``maxit`` iterations."""
        if self.itsolver is None:
            raise NotImplementedError("ItSolver is an abstract class")
        if K is None:
            K = self.precon
        x[:] = 0.0
        if K is None:
            info, it, relres = self.itsolver(self.matrix, rhs, x, tol, maxit)
        else:
            info, it, relres = self.itsolver(self.matrix, rhs, x, tol, maxit, K)
        self.nofCalled += 1
        self.totalIterations += it
        self.lastIterations = self.itn = it
        self.lastInfo = info
        self.resNorm = relres
        self.converged = info >= 0
        if info < 0 and self.debug:
            raise RuntimeError("%s: info = %d after %d iterations (relres %g)" % (self.name, info, it, relres))
        return info, it, relres

    def __call__(self, *args, **kwargs):
        return self.solve(*args, **kwargs)

    def __repr__(self):
        return "<%s solver, called %d times, %d iterations>" % (self.name, self.nofCalled, self.totalIterations)


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
