"""JDSYM -- Jacobi-Davidson eigensolver for symmetric (generalised) eigenproblems
``A x = lambda M x`` (drop-in for ``pysparse.eigen.jdsym``).

Restatement of the reference's algorithm (``pysparse/eigen/src/jdsym.c:52-652``, the
correction-equation operator ``correq.c:28-262`` and the orthogonalisation routines
``orthopack.c:38-155``) with the reference's call signature, keyword names, defaults and
return tuple ``(kconv, lambda, Q, it, it_inner)`` (``jdsymmodule.c:178-207,300``).

Division of labour, as in the reference: the products with ``A`` and ``M``, the preconditioner
and the inner iterative solver (``linsolver``, e.g. ``krylov.qmrs``) are the GPU path -- the
correction equation reaches the solver as a duck-typed operator object with ``shape``,
``matvec`` and ``precon``, exactly what ``ItSolvers_Solve`` (``spmatrixmodule.c:266-307``)
hands to it in the reference; the outer iteration (Rayleigh-Ritz on a search space of at most
``jmax`` vectors, Gram-Schmidt, restarts, bookkeeping) is small dense algebra on the host
(LAPACK/BLAS there, NumPy here).

When ``A``, ``M`` and ``K`` are objects of this package the correction equation itself is a
device object as well (``pysparse_b200.eigen.correq``, ``csrc/psp_correq.cu``): the blocks
``Q``, ``M Q`` and ``Y`` are mirrored in HBM -- only the columns that changed are uploaded
before a correction equation -- and the operator exposes the ``_psp_mat`` / ``_psp_prec`` handles
by which ``krylov.*`` recognises a native operand, so an inner solve runs without a single
vector crossing PCIe.  With foreign operators (anything that only has ``matvec`` / ``precon``)
the operator below does the same arithmetic in NumPy around their callbacks.
"""
import numpy as np

from . import correq as _device_correq

_DEVICE_KMAX = 64            # kCeqMaxK of csrc/psp_correq.cuh

__all__ = ["jdsym", "CorrEq", "JDSYM_OP_UNSYM", "JDSYM_OP_SYM"]

JDSYM_OP_UNSYM = 1
JDSYM_OP_SYM = 2

#: False: keep the correction equation on the host-callback protocol even for device operands
#: (parity tests compare the two)
DEVICE_CORREQ = True


# ---------------------------------------------------------------- operator helpers
def _order(op, what):
    shape = getattr(op, "shape", None)
    if shape is None or len(shape) != 2 or shape[0] != shape[1]:
        raise ValueError("%s must have a square shape attribute" % what)
    return int(shape[0])


def _apply(method, x):
    y = np.empty_like(x)
    method(x, y)
    return y


# ---------------------------------------------------------------- orthopack.c
def _mgs(u, m, Q):
    "modified Gram-Schmidt: u := (I - Q_m Q_m') u, one column at a time (orthopack.c:118-127)"
    for i in range(m):
        u -= np.dot(Q[:, i], u) * Q[:, i]


def _mgsm(u, m, Q, QM):
    "M-orthogonal modified Gram-Schmidt (orthopack.c:139-148)"
    for i in range(m):
        u -= np.dot(QM[:, i], u) * Q[:, i]


def _icgs(u, m, A):
    "iterated classical Gram-Schmidt against A[:, :m]; returns ||u|| (orthopack.c:38-64)"
    nrm = float(np.linalg.norm(u))
    if m == 0:
        return nrm
    for _ in range(5):
        u -= A[:, :m] @ (A[:, :m].T @ u)
        old, nrm = nrm, float(np.linalg.norm(u))
        if nrm > 0.5 * old:
            break
    return nrm


def _icgsm(u, m, Q, mmat):
    "iterated classical M-orthogonal Gram-Schmidt; returns the M-norm of u (orthopack.c:74-107)"
    um = _apply(mmat.matvec, u)
    nrm = float(np.sqrt(np.dot(u, um)))
    if m == 0:
        return nrm
    for _ in range(5):
        u -= Q[:, :m] @ (Q[:, :m].T @ um)
        um = _apply(mmat.matvec, u)
        old, nrm = nrm, float(np.sqrt(np.dot(u, um)))
        if nrm > 0.5 * old:
            break
    return nrm


# ---------------------------------------------------------------- correq.c
class CorrEq:
    """The correction-equation operator handed to the inner solver as both ``A`` and ``K``:
    ``shape``, ``matvec(x, y)``, ``precon(x, y)`` (correq.c:144-262).  ``update(k, theta)``
    selects the number of locked vectors and the shift, ``right(r)`` projects a residual
    into the right-hand side in place."""

    def __init__(self, optype, n, amat, mmat, prec, Q, Qm, Y, H, device=True):
        self.optype, self.n = optype, n
        self.amat, self.mmat, self.prec = amat, mmat, prec
        self.Q, self.Qm, self.Y, self.H = Q, Qm, Y, H
        self.shape = (n, n)
        self.k = 0
        self.theta = 0.0
        self._dev = None
        self._synced = 0                             # columns below this index are up to date on the device
        if device and n > 0 and Q.shape[1] <= _DEVICE_KMAX:
            try:
                self._dev = _device_correq.operator(optype, amat, mmat, prec, Q.shape[1])
            except TypeError:                        # a foreign operand: stay on the callback protocol
                self._dev = None

    @property
    def _psp_mat(self):
        if self._dev is None:
            raise AttributeError("_psp_mat")
        return self._dev._psp_mat

    @property
    def _psp_prec(self):
        if self._dev is None:
            raise AttributeError("_psp_prec")
        return self._dev._psp_prec

    def update(self, k, theta, stable=0):
        """Use the first ``k`` columns and the shift ``theta``.  ``stable``: columns below this
        index have not been written since the previous call (the locked eigenvectors)."""
        self.k, self.theta = k, theta
        if self._dev is None:
            return
        lo = min(self._synced, stable, k)
        if lo < k:
            self._dev.set_columns(0, lo, self.Q[:, lo:k])
            if self.mmat is not None:
                self._dev.set_columns(1, lo, self.Qm[:, lo:k])
            if self.prec is not None:
                self._dev.set_columns(2, lo, self.Y[:, lo:k])
        self._synced = min(stable, k)
        hinv = np.linalg.inv(self.H[:k, :k]) if (self.prec is not None and k > 0) else None
        self._dev.update(k, float(theta), hinv)

    # y := (A - theta M) x   /   (A - theta I) x
    def _a_theta(self, x):
        y = _apply(self.amat.matvec, x)
        if self.mmat is not None:
            if self.theta != 0.0:
                y -= self.theta * _apply(self.mmat.matvec, x)
        else:
            y -= self.theta * x
        return y

    def _project1(self, A, B, x):
        "(I - A B') x"
        k = self.k
        return x - A[:, :k] @ (B[:, :k].T @ x)

    def _project2(self, Qx, x):
        "(I - Y inv(H) Qx') x"
        k = self.k
        w = np.linalg.solve(self.H[:k, :k], Qx[:, :k].T @ x)
        return x - self.Y[:, :k] @ w

    def right(self, r):
        if self.optype == JDSYM_OP_SYM:
            if self.mmat is not None:
                r[:] = self._project1(self.Qm, self.Q, r)
            else:
                _mgs(r, self.k, self.Q)
        elif self.prec is not None:
            Qx = self.Qm if self.mmat is not None else self.Q
            r[:] = self._project2(Qx, _apply(self.prec.precon, r))
        elif self.mmat is not None:
            r[:] = self._project1(self.Q, self.Qm, r)
        else:
            _mgs(r, self.k, self.Q)

    def matvec(self, x, y):
        if self._dev is not None:
            return self._dev.matvec(x, y)
        x = np.ascontiguousarray(x, dtype=np.float64)
        t = self._a_theta(x)
        if self.optype == JDSYM_OP_SYM:
            y[:] = self._project1(self.Qm, self.Q, t) if self.mmat is not None else self._project1(self.Q, self.Q, t)
        elif self.prec is not None:
            Qx = self.Qm if self.mmat is not None else self.Q
            y[:] = self._project2(Qx, _apply(self.prec.precon, t))
        elif self.mmat is not None:
            y[:] = self._project1(self.Q, self.Qm, t)
        else:
            y[:] = self._project1(self.Q, self.Q, t)

    def precon(self, x, y):
        if self._dev is not None:
            return self._dev.precon(x, y)
        x = np.ascontiguousarray(x, dtype=np.float64)
        if self.optype == JDSYM_OP_SYM:
            if self.prec is not None:
                Qx = self.Qm if self.mmat is not None else self.Q
                y[:] = self._project2(Qx, _apply(self.prec.precon, x))
            elif self.mmat is not None:
                y[:] = self._project1(self.Q, self.Qm, x)
            else:
                y[:] = x
        else:                                        # the preconditioner is inside the operator
            y[:] = x


# ---------------------------------------------------------------- jdsym.c
def _start_vectors(count, n):
    "uniform (0, 1) start vectors: LAPACK dlarnv(1, seed (2,3,5,7)) like jdsym.c:104-105,228-232"
    try:
        from scipy.linalg.lapack import dlarnv
        return np.asarray(dlarnv(1, np.array([2, 3, 5, 7], dtype=np.int32), count * n)[0]).reshape((n, count), order="F")
    except Exception:                                # pragma: no cover  (SciPy missing)
        return np.random.RandomState(2357).random_sample((n, count))


def _sorteig(s, U, tau, strategy):
    "Ritz pairs by distance to tau; strategy 1 sends values below tau to the end (jdsym.c:769-812)"
    key = np.abs(s - tau)
    if strategy == 1:
        key = np.where(s < tau, np.inf, key)
    idx = np.argsort(key, kind="stable")
    return s[idx], U[:, idx]


def _print_status(clvl, it, k, j, blksize, actblksize, s, resnrm, actcorrits):
    if clvl < 1:
        return
    if blksize == 1:
        if it == 0:
            print("  IT   K   J       RES CGIT RITZVALS(1:5)")
            print("-" * 77)
        print("%4d %3d %3d %9.2e %4d" % (it + 1, k, j, resnrm[0], actcorrits[0]) +
              "".join(" %9.2e" % v for v in s[:min(j, 5)]))
    else:
        print("%4d %3d %3d" % (it + 1, k, j) + "".join(" %9.2e" % v for v in s[:actblksize]) + "  " +
              "".join(" %9.2e" % v for v in resnrm[:actblksize]) + "  " +
              "".join(" %4d" % v for v in actcorrits[:actblksize]))


def jdsym(A, M, K, kmax, tau, jdtol, itmax, linsolver, jmax=25, jmin=10, blksize=1, blkwise=0, V0=None,
          optype=JDSYM_OP_SYM, linitmax=200, eps_tr=1.0e-3, toldecay=1.5, clvl=0, strategy=0, projector=None):
    """jdsym(A, M, K, kmax, tau, jdtol, itmax, linsolver, **kwargs)

    Jacobi-Davidson method for ``kmax`` eigenpairs of ``A x = lambda M x`` (``M=None``:
    ``A x = lambda x``) near the target ``tau``.  ``A`` and ``M`` provide ``shape`` and
    ``matvec``; ``K`` (or ``None``) ``shape`` and ``precon``; ``linsolver`` is one of the
    ``krylov`` solvers (or any function with their signature and return tuple).  Keywords:
    ``jmax jmin blksize blkwise V0 optype linitmax eps_tr toldecay clvl strategy projector``
    (doc/pysparse/source/jdsym.rst).  Returns ``(kconv, lambda, Q, it, it_inner)``:
    number of converged pairs, their eigenvalues, eigenvectors ``Q[:, i]``, outer iterations,
    total inner iterations."""
    n = _order(A, "A")
    for op, what in ((M, "M"), (K, "K"), (projector, "projector")):
        if op is not None and _order(op, what) != n:
            raise ValueError("matrix, preconditioner or projector shapes differ")
    kmax, itmax, jmax, jmin, blksize = int(kmax), int(itmax), int(jmax), int(jmin), int(blksize)
    # the reference only asserts these (compiled out in release builds); a search space
    # cannot usefully exceed the dimension, so tiny problems get their limits clamped
    jmax = min(jmax, n)
    jmin = max(1, min(jmin, jmax - 1)) if jmax > 1 else 1
    if not (jdtol > 0 and 0 < kmax <= n and itmax >= 0 and linitmax >= 0 and eps_tr >= 0.0 and toldecay > 1.0):
        raise ValueError("jdsym: invalid parameter (need jdtol > 0, 0 < kmax <= n, toldecay > 1, ...)")
    if not (0 < blksize <= kmax and blksize <= jmin and (blksize <= jmax - jmin or jmax == 1)):
        raise ValueError("jdsym: blksize must satisfy 0 < blksize <= kmax, blksize <= jmin, blksize <= jmax - jmin")
    if blkwise not in (0, 1) or optype not in (JDSYM_OP_UNSYM, JDSYM_OP_SYM) or strategy not in (0, 1):
        raise ValueError("jdsym: blkwise must be 0/1, optype 1/2, strategy 0/1")
    if V0 is not None:
        V0 = np.asarray(V0)
        if V0.dtype != np.float64 or V0.ndim not in (1, 2) or V0.shape[0] != n:
            raise ValueError("V0 is not of correct type or shape")

    if clvl >= 1:
        print("JDSYM     Solving  %s  %s preconditioning.\n" % ("A*x = lambda*x" if M is None else "A*x = lambda*M*x",
                                                                "without" if K is None else "with"))
        print("  N=      %10d  ITMAX=%4d" % (n, itmax))
        print("  KMAX=%3d  JMIN=%3d  JMAX=%3d  V0DIM=%3d" % (kmax, jmin, jmax, 0 if V0 is None else (1 if V0.ndim == 1 else V0.shape[1])))
        print("  BLKSIZE=        %2d  BLKWISE=      %5s" % (blksize, "TRUE" if blkwise else "FALSE"))
        print("  TAU=   %11.4e  JDTOL=  %11.4e  STRATEGY= %8d" % (tau, jdtol, strategy))
        print("  OPTYPE=%12s" % ("UNSYM" if optype == JDSYM_OP_UNSYM else "SYM"))
        print("  LINITMAX=    %5d  EPS_TR=  %10.3e  TOLDECAY=%9.2e\n" % (linitmax, eps_tr, toldecay))

    V = np.zeros((n, jmax + blksize), order="F")        # search space (one spare block: clamped jmax)
    Q = np.zeros((n, kmax), order="F")                  # locked eigenvectors, then the active Ritz vectors
    Qm = np.zeros((n, kmax), order="F") if M is not None else None
    Y = np.zeros((n, kmax), order="F") if K is not None else None
    H = np.zeros((kmax, kmax)) if K is not None else None
    Res = np.zeros((n, blksize), order="F")
    resnrm = np.zeros(blksize)
    lam = np.zeros(kmax)
    Mp = np.zeros((jmax + blksize, jmax + blksize))     # interaction matrix V'AV, upper triangle
    correq = CorrEq(optype, n, A, M, K, Q, Qm, Y, H, device=DEVICE_CORREQ)
    project = (lambda v: projector.project(v)) if projector is not None else None

    # ---- initial search space (jdsym.c:212-262)
    j = 0
    if V0 is not None:
        if V0.ndim == 1:
            V[:, 0] = V0; j = 1
        else:
            j = min(V0.shape[1], jmax)
            V[:, :j] = V0[:, :j]
    if j < blksize:
        V[:, j:blksize] = _start_vectors(blksize - j, n)
        j = blksize
    for c in range(j):
        v = np.ascontiguousarray(V[:, c])
        if project:
            project(v)
        if M is None:
            _mgs(v, c, V)
            v /= np.linalg.norm(v)
        else:
            v /= _icgsm(v, c, V, M)
        V[:, c] = v
    for c in range(j):
        Mp[:c + 1, c] = V[:, :c + 1].T @ _apply(A.matvec, np.ascontiguousarray(V[:, c]))

    k = 0
    it = 0
    nof_ce = nof_ce_its = 0
    actblksize = blksize
    solvestep = [1] * blksize
    actcorrits = [0] * blksize
    s = np.zeros(0)
    done = False

    while it < itmax and not done:
        # ---- projected eigenproblem (jdsym.c:286-298)
        T = np.triu(Mp[:j, :j])
        s, U = np.linalg.eigh(T + np.triu(T, 1).T)
        s, U = _sorteig(s, U, tau, strategy)

        found = True
        while found:
            convind, keepind = [], []
            for act in range(actblksize):
                q = V[:, :j] @ U[:j, act]
                Q[:, k + act] = q
                aq = _apply(A.matvec, q)
                if M is None:
                    r = aq - s[act] * q
                    vec = q
                else:
                    qm = _apply(M.matvec, q)
                    Qm[:, k + act] = qm
                    r = aq - s[act] * qm
                    vec = qm
                if K is not None:
                    mat = Qm if M is not None else Q
                    y = _apply(K.precon, np.ascontiguousarray(vec))
                    Y[:, k + act] = y
                    H[:k + act + 1, k + act] = mat[:, :k + act + 1].T @ y
                    H[k + act, :k + act + 1] = Y[:, :k + act + 1].T @ vec
                Res[:, act] = r
                resnrm[act] = np.linalg.norm(r)
                (convind if resnrm[act] < jdtol else keepind).append(act)
            conv, keep = len(convind), len(keepind)
            found = ((blkwise == 1 and conv == actblksize) or (blkwise == 0 and conv != 0)) and \
                    (j > actblksize or k == kmax - actblksize)

            if found:                                               # jdsym.c:398-489
                for a, ci in enumerate(convind):
                    lam[k + a] = s[ci]
                s = np.concatenate([s[keepind], s[actblksize:j]])
                rest = V[:, :j] @ U[:j, actblksize:j]               # the Ritz vectors not looked at yet
                kept = Q[:, [k + ki for ki in keepind]].copy()
                V[:, :keep] = kept
                V[:, keep:keep + rest.shape[1]] = rest
                for a, ci in enumerate(convind):                    # lock the converged ones (ci >= a)
                    Q[:, k + a] = Q[:, k + ci]
                    if M is not None:
                        Qm[:, k + a] = Qm[:, k + ci]
                    if K is not None:
                        Y[:, k + a] = Y[:, k + ci]
                        d = k + a
                        H[:d, d] = H[:d, k + ci]
                        H[d, d] = H[k + ci, k + ci]
                        H[d, :d] = H[k + ci, :d]
                j -= conv
                Mp[:, :] = 0.0
                Mp[np.arange(j), np.arange(j)] = s[:j]
                U = np.eye(j)
                if strategy == 1:
                    tau = max([tau] + [lam[k + a] for a in range(conv)])
                k += conv
                actblksize = min(blksize, kmax - k)
                if k == kmax:
                    done = True
                    break
                solvestep = [solvestep[ki] for ki in keepind] + [1] * (blksize - keep)

            if j + actblksize > jmax:                               # restart (jdsym.c:499-513)
                V[:, :jmin] = V[:, :j] @ U[:j, :jmin]
                j = jmin
                U = np.eye(j)
                Mp[:, :] = 0.0
                Mp[np.arange(j), np.arange(j)] = s[:j]
        if done:
            break

        # ---- correction equations (jdsym.c:525-603)
        for act in range(actblksize):
            v = np.zeros(n)
            r = np.ascontiguousarray(Res[:, act])
            correq.update(k + actblksize, float(s[act]) if resnrm[act] < eps_tr else tau, stable=k)
            it_tol = toldecay ** (-solvestep[act])
            solvestep[act] += 1
            correq.right(r)
            res = linsolver(correq, r, v, it_tol, linitmax, correq)
            linit = int(res[1])
            nof_ce += 1
            nof_ce_its += linit
            actcorrits[act] = linit
            if M is not None:
                _mgsm(v, k + actblksize, Q, Qm)
                if project:
                    project(v)
                alpha = _icgsm(v, j, V, M)
            else:
                _mgs(v, k + actblksize, Q)
                if project:
                    project(v)
                alpha = _icgs(v, j, V)
            v /= alpha
            V[:, j] = v
            Mp[:j + 1, j] = V[:, :j + 1].T @ _apply(A.matvec, v)
            j += 1
        _print_status(clvl, it, k, j - blksize, blksize, actblksize, s, resnrm, actcorrits)
        it += 1

    if clvl >= 1:
        print("\nJDSYM execution statistics\n")
        print("IT_OUTER=%d   IT_INNER_TOT=%d   IT_INNER_AVG=%8.2f   IT_INNER_PER_OUTER=%8.2f" %
              (it, nof_ce_its, nof_ce_its / max(it, 1), nof_ce_its / max(nof_ce, 1)))
        print("\nConverged eigensolutions in order of convergence:\n")
        print("  I              LAMBDA(I)      RES(I)")
        print("---------------------------------------")
        for a in range(k):
            q = np.ascontiguousarray(Q[:, a])
            r = _apply(A.matvec, q) - lam[a] * (q if M is None else _apply(M.matvec, q))
            print("%3d %22.15e %12.5e" % (a + 1, lam[a], np.linalg.norm(r)))
        print()
    return k, lam[:k].copy(), np.ascontiguousarray(Q[:, :k]), it, nof_ce_its
