"""Operator-style matrix classes on top of ``ll_mat`` (Python 3).

Provides what ``pysparse.sparse.pysparseMatrix`` provides in the reference
(``pysparse/sparse/pysparseMatrix.py``, ``sparseMatrix.py``; documented in
``doc/pysparse/source/pysparsematrix.rst``): ``PysparseMatrix`` owns an ``ll_mat`` and lets it
take part in expressions -- ``A + B``, ``A - B``, ``A * x``, ``x * A``, ``A * B``, ``2 * A``,
``-A`` -- plus the bulk accessors ``put / take / addAt / find`` and their diagonal variants;
``PysparseIdentityMatrix`` and ``PysparseSpDiagsMatrix`` are ready-made constructors.  The
reference's worked examples for these classes are kept as tests
(``tests/test_host_pysparse_matrix.py``).

Matrix-vector products run on the GPU.  The wrapped ``ll_mat`` keeps a device copy keyed on a
modification counter, so a loop of ``A * x`` uploads the matrix once and each product moves
only the two vectors.  There is no CPU path.
"""
import numpy

from . import spmatrix

__docformat__ = "restructuredtext"
__all__ = ["SparseMatrix", "PysparseMatrix", "PysparseIdentityMatrix", "PysparseSpDiagsMatrix",
           "PysparseMatrix4Scipy"]

_CELL = 11          # width of one printed cell


def _is_number(x):
    return isinstance(x, (int, float, numpy.integer, numpy.floating)) and not isinstance(x, bool)


def _cell(v):
    """One cell of the printed matrix: '---' for a zero, fixed point while the magnitude is
    within e^-4 .. e^4, two-digit scientific notation beyond."""
    if v == 0:
        return "---".center(_CELL)
    text = "%9.6f" % v if abs(numpy.log(abs(v))) <= 4 else "%9.2e" % v
    return text.ljust(_CELL)


class SparseMatrix:
    """Common behaviour of the matrix classes: text rendering and the operator identities
    that follow from ``__add__`` / ``__sub__`` / ``__mul__``."""

    __array_priority__ = 100.0
    __array_ufunc__ = None          # ``ndarray * A`` is handed to ``A.__rmul__`` as a whole

    def getShape(self):
        raise NotImplementedError("abstract class")

    def __str__(self):
        nrow, ncol = self.getShape()
        return "\n".join("".join(_cell(self[i, j]) for j in range(ncol)) for i in range(nrow))

    def __repr__(self):
        return str(self.getMatrix())

    def __radd__(self, other):
        return self + other

    def __rsub__(self, other):
        return -(self - other)

    def __neg__(self):
        return self * -1

    def __pos__(self):
        return self


class PysparseMatrix(SparseMatrix):
    """A sparse matrix that can be populated conveniently and used in expressions.

    Keyword arguments (all optional):

    ``nrow``, ``ncol``  shape of a new, empty matrix;  ``size``  shorthand for a square one;
    ``symmetric``       store only the lower triangle (square matrices);
    ``bandwidth``       expected bandwidth, ``sizeHint``  expected number of entries (only
                        used to pre-size the container);
    ``storeZeros``      keep explicitly assigned zeros;
    ``matrix``          wrap this existing ``ll_mat`` instead of creating one.
    """

    def __init__(self, *, nrow=0, ncol=0, size=0, bandwidth=0, matrix=None, sizeHint=0,
                 symmetric=False, storeZeros=False):
        if matrix is not None:
            self.matrix = matrix
            return
        if size > 0:
            if (nrow > 0 or ncol > 0) and not (nrow == size and ncol == size):
                raise ValueError("size argument was given but does not match nrow and ncol")
            nrow = ncol = size
        symmetric = bool(symmetric) and nrow == ncol
        hint = int(sizeHint or 0)
        if hint <= 0:                              # a rough count from the band structure
            hint = nrow if symmetric else min(nrow, ncol)
            if bandwidth > 0:
                hint = bandwidth * max(hint, 1)
        hint = max(hint, 1)
        if symmetric:
            self.matrix = spmatrix.ll_mat_sym(nrow, hint, int(bool(storeZeros)))
        else:
            self.matrix = spmatrix.ll_mat(nrow, ncol, hint, int(bool(storeZeros)))

    # ---- queries
    def isSymmetric(self):
        "``True`` when only the lower triangle is stored."
        return bool(self.matrix.issym)

    def getMatrix(self):
        "The wrapped ``ll_mat``."
        return self.matrix

    def getShape(self):
        "``(nrow, ncol)``."
        return self.matrix.shape

    def getNnz(self):
        "Number of stored entries."
        return self.matrix.nnz

    shape = property(getShape)
    nnz = property(getNnz)

    def copy(self):
        "A deep copy."
        return PysparseMatrix(matrix=self.matrix.copy())

    # ---- indexing: everything ``ll_mat`` accepts (integers, slices, lists, index arrays)
    def __getitem__(self, index):
        item = self.matrix[index]
        return item if isinstance(item, (int, float)) else PysparseMatrix(matrix=item)

    def __setitem__(self, index, value):
        self.matrix[index] = value.getMatrix() if isinstance(value, PysparseMatrix) else value

    # ---- sums
    def _shifted(self, sign, other, inplace):
        if self.getShape() != other.getShape():
            raise TypeError("Only sparse matrices of the same size may be %s" % ("added" if sign > 0 else "subtracted"))
        target = self.matrix if inplace else self.matrix.copy()
        if target.issym and not other.isSymmetric():
            target.generalize()                    # symmetric (+/-) general needs general storage
        target.shift(sign, other.getMatrix())
        return self if inplace else PysparseMatrix(matrix=target)

    def __add__(self, other):
        """``A + B`` for two matrices of equal shape; ``A + c`` adds the number ``c`` to every
        STORED entry (``A + 0`` is ``A`` itself)."""
        if _is_number(other):
            if other == 0:
                return self
            result = self.copy()
            val, irow, jcol = result.find()
            result.matrix.update_add_at(numpy.full(len(val), float(other)), irow, jcol)
            return result
        if isinstance(other, PysparseMatrix):
            return self._shifted(+1, other, inplace=False)
        return NotImplemented

    def __sub__(self, other):
        if _is_number(other):
            return self + (-other)
        if isinstance(other, PysparseMatrix):
            return self._shifted(-1, other, inplace=False)
        return NotImplemented

    def __iadd__(self, other):
        return self._shifted(+1, other, inplace=True)

    def __isub__(self, other):
        return self._shifted(-1, other, inplace=True)

    # ---- products
    def __mul__(self, other):
        """``A * B`` (sparse product, a new matrix), ``A * x`` for a vector of length ``ncol``
        (a new NumPy array, computed on the GPU) or ``A * c`` for a number."""
        nrow, ncol = self.getShape()
        if isinstance(other, PysparseMatrix):
            if other.getShape()[0] != ncol:
                raise TypeError("Matrices dimensions do not match for product")
            return PysparseMatrix(matrix=spmatrix.matrixmultiply(self.matrix, other.getMatrix()))
        shape = numpy.shape(other)
        if shape == ():
            scaled = self.matrix.copy()
            scaled.scale(float(other))
            return PysparseMatrix(matrix=scaled)
        if shape == (ncol,):
            y = numpy.empty(nrow)
            self.matrix.matvec(numpy.ascontiguousarray(other, dtype=numpy.float64), y)
            return y
        raise TypeError("Cannot multiply objects")

    def __rmul__(self, other):
        "``x * A`` for a vector of length ``nrow`` is ``A^T x``; ``c * A`` scales."
        if isinstance(other, numpy.ndarray) and other.ndim == 1:
            nrow, ncol = self.getShape()
            if other.shape != (nrow,):
                raise TypeError("Cannot multiply objects")
            y = numpy.empty(ncol)
            self.matrix.matvec_transp(numpy.ascontiguousarray(other, dtype=numpy.float64), y)
            return y
        return self * other

    def __imul__(self, other):
        if not _is_number(other):
            raise TypeError("In-place multiplication is with scalars only")
        self.matrix.scale(float(other))
        return self

    def matvec(self, x):
        "``A * x`` (the one-argument form SciPy-style solvers call)."
        return self * x

    def col_scale(self, v):
        "``A[:, j] *= v[j]`` in place."
        return self.matrix.col_scale(v)

    def row_scale(self, v):
        "``A[i, :] *= v[i]`` in place."
        return self.matrix.row_scale(v)

    # ---- bulk access
    def find(self):
        "``(val, irow, jcol)``: the stored entries in coordinate format, row by row."
        return self.matrix.find()

    def put(self, value, id1=None, id2=None):
        """``A[id1[k], id2[k]] = value[k]``.  A scalar ``value`` stands for a constant vector.
        Missing ``id2`` means ``id1``; missing both means the leading diagonal positions
        (``range(len(value))``, or the whole diagonal for a scalar)."""
        nrow, ncol = self.getShape()
        if id1 is not None and id2 is None:
            id2 = id1
        count = None if _is_number(value) else len(value)
        if id1 is None:
            id1 = numpy.arange(nrow if count is None else count, dtype=numpy.int64)
        if id2 is None:
            id2 = numpy.arange(ncol if count is None else count, dtype=numpy.int64)
        self.matrix.put(value, numpy.asarray(id1, dtype=numpy.int64), numpy.asarray(id2, dtype=numpy.int64))

    def putDiagonal(self, vector):
        "Assign the leading ``len(vector)`` diagonal entries (all of them for a scalar)."
        if _is_number(vector):
            ids = numpy.arange(self.getShape()[0])
            self.put(vector, ids, ids)
        else:
            self.matrix.put(vector)

    def take(self, id1=None, id2=None):
        "A new array with ``A[id1[k], id2[k]]`` (the diagonal when no indices are given)."
        if id1 is None:
            id1 = numpy.arange(min(self.getShape()), dtype=numpy.int64)
        if id2 is None:
            id2 = id1
        out = numpy.zeros(len(id1))
        self.matrix.take(out, numpy.asarray(id1, dtype=numpy.int64), numpy.asarray(id2, dtype=numpy.int64))
        return out

    def takeDiagonal(self):
        "The diagonal as a NumPy array."
        return self.take()

    def addAt(self, vector, id1, id2):
        "``A[id1[k], id2[k]] += vector[k]``; repeated positions accumulate."
        self.matrix.update_add_at(vector, id1, id2)

    def addAtDiagonal(self, vector):
        "Add a vector (to the leading diagonal entries) or a number (to the whole diagonal)."
        if _is_number(vector):
            vector = numpy.full(self.getShape()[0], float(vector))
        ids = numpy.arange(len(vector))
        self.addAt(vector, ids, ids)

    def getNumpyArray(self):
        "A dense copy."
        nrow, ncol = self.getShape()
        rows, cols = numpy.indices((nrow, ncol))
        return self.take(rows.ravel(), cols.ravel()).reshape(nrow, ncol)

    def exportMmf(self, filename):
        "Write the matrix in MatrixMarket coordinate format."
        self.matrix.export_mtx(filename)


class PysparseIdentityMatrix(PysparseMatrix):
    "The ``size`` x ``size`` identity, stored symmetrically."

    def __init__(self, size):
        super().__init__(size=size, bandwidth=1, symmetric=True)
        self.putDiagonal(numpy.ones(size))


class PysparseSpDiagsMatrix(PysparseMatrix):
    """A square banded matrix given by its diagonals: ``vals[k]`` goes on diagonal ``pos[k]``
    (0 the main diagonal, positive above, negative below); diagonal ``d`` has ``size - |d|``
    entries and takes that many leading elements of ``vals[k]``.  With ``symmetric=True`` the
    diagonals above the main one are not stored."""

    def __init__(self, size, vals, pos, **kwargs):
        pos = [int(d) for d in pos]
        kwargs.pop("bandwidth", None)
        kwargs.pop("sizeHint", None)
        super().__init__(size=size, bandwidth=max([abs(d) for d in pos] + [0]),
                         sizeHint=sum(size - abs(d) for d in pos), **kwargs)
        for d, v in zip(pos, vals):
            length = size - abs(d)
            if length <= 0 or (d > 0 and self.isSymmetric()):
                continue
            k = numpy.arange(length)
            rows, cols = (k, k + d) if d >= 0 else (k - d, k)
            self.put(numpy.asarray(v, dtype=numpy.float64)[:length], rows, cols)


class PysparseMatrix4Scipy(PysparseMatrix):
    "A ``PysparseMatrix`` whose ``matvec(x, y)`` has the two-argument ``ll_mat`` form."

    def matvec(self, x, y):
        return self.matrix.matvec(x, y)
