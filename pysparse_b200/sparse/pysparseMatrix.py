"""Higher-level sparse matrix classes over ``ll_mat`` (Python 3).

Drop-in for ``pysparse.sparse.pysparseMatrix`` (reference:
``pysparse/sparse/pysparseMatrix.py`` and ``sparseMatrix.py``; documented in
``doc/pysparse/source/pysparsematrix.rst``): ``PysparseMatrix`` wraps an ``ll_mat`` and adds
the arithmetic operators (``A + B``, ``A - B``, ``A * x``, ``x * A``, ``A * B``, ``2 * A``),
``PysparseIdentityMatrix`` and ``PysparseSpDiagsMatrix`` build common matrices.

The wrapped ``ll_mat`` keeps a device copy that is rebuilt only when the matrix has been
modified, so ``A * x`` in a loop uploads the matrix once and each product moves only the two
vectors (the products themselves run on the GPU; there is no CPU path).
"""
import numpy

from . import spmatrix

__docformat__ = "restructuredtext"
__all__ = ["SparseMatrix", "PysparseMatrix", "PysparseIdentityMatrix", "PysparseSpDiagsMatrix",
           "PysparseMatrix4Scipy"]


def _is_scalar(x):
    return isinstance(x, (int, float, numpy.integer, numpy.floating)) and not isinstance(x, bool)


class SparseMatrix:
    """Abstract base: printing and the operator identities shared by all matrix classes
    (reference ``sparseMatrix.py:49-140``)."""

    __array_priority__ = 100.0          # numpy defers ``ndarray * A`` to ``A.__rmul__``
    __array_ufunc__ = None

    def getShape(self):
        raise NotImplementedError

    def __str__(self):
        # the reference's cell format (sparseMatrix.py:75-95): 11-wide cells, '---' for zeros,
        # %9.6f while |ln|v|| <= 4, %9.2e beyond
        width = 11
        nrow, ncol = self.getShape()
        out = []
        for i in range(nrow):
            row = ""
            for j in range(ncol):
                v = self[i, j]
                if v == 0:
                    row += "---".center(width)
                elif abs(numpy.log(abs(v))) <= 4:
                    row += ("%9.6f" % v).ljust(width)
                else:
                    row += ("%9.2e" % v).ljust(width)
            out.append(row)
        return "\n".join(out)

    def __repr__(self):
        return str(self.getMatrix())

    def __radd__(self, other):
        return self.__add__(other)

    def __rsub__(self, other):
        return -(self.__sub__(other))

    def __neg__(self):
        return self * -1

    def __pos__(self):
        return self


class PysparseMatrix(SparseMatrix):
    """
    A class wrapper for the ``ll_mat`` sparse matrix type that makes populating matrices
    and operating on matrices and vectors read naturally.

    :keywords:
        :nrow:       The number of rows of the matrix
        :ncol:       The number of columns of the matrix
        :size:       The common number of rows and columns, for a square matrix
        :bandwidth:  The bandwidth (if creating a band matrix)
        :matrix:     The starting ``ll_mat`` if there is one
        :sizeHint:   A guess on the number of nonzero elements of the matrix
        :symmetric:  A boolean indicating whether the matrix is symmetric.
        :storeZeros: A boolean indicating whether to store zero values.
    """

    def __init__(self, **kwargs):
        nrow = kwargs.get("nrow", 0)
        ncol = kwargs.get("ncol", 0)
        bandwidth = kwargs.get("bandwidth", 0)
        matrix = kwargs.get("matrix", None)
        sizeHint = kwargs.get("sizeHint", 0)
        storeZeros = bool(kwargs.get("storeZeros", False))
        symmetric = bool(kwargs.get("symmetric", False))
        size = kwargs.get("size", 0)

        if size > 0:
            if nrow > 0 or ncol > 0:
                if size != nrow or size != ncol:
                    raise ValueError("size argument was given but does not match nrow and ncol")
            else:
                nrow = ncol = size

        if matrix is not None:
            self.matrix = matrix
            return
        if symmetric and nrow == ncol:
            if not sizeHint:
                sizeHint = nrow
                if bandwidth > 0:
                    sizeHint += 2 * (bandwidth - 1) * (2 * nrow - bandwidth - 2)
            self.matrix = spmatrix.ll_mat_sym(nrow, max(int(sizeHint), 1), int(storeZeros))
        else:
            if not sizeHint:
                sizeHint = min(nrow, ncol)
                if bandwidth > 0:
                    sizeHint = bandwidth * (2 * sizeHint - bandwidth - 1) // 2
            self.matrix = spmatrix.ll_mat(nrow, ncol, max(int(sizeHint), 1), int(storeZeros))

    # ---- queries ---------------------------------------------------------------------
    def isSymmetric(self):
        "Returns `True` is `self` is a symmetric matrix or `False` otherwise"
        return bool(self.matrix.issym)

    def getNnz(self):
        "Returns the number of nonzero elements of `self`"
        return self.matrix.nnz

    def getMatrix(self):
        "Returns the underlying `ll_mat` sparse matrix of `self`"
        return self.matrix

    def getShape(self):
        "Returns the shape ``(nrow,ncol)`` of a sparse matrix"
        return self.matrix.shape

    nnz = property(getNnz)
    shape = property(getShape)

    def copy(self):
        "Returns a (deep) copy of a sparse matrix"
        return PysparseMatrix(matrix=self.matrix.copy())

    # ---- element access (fancy indexing carries over from ll_mat) ---------------------
    def __getitem__(self, index):
        m = self.matrix[index]
        if isinstance(m, (int, float)):
            return m
        return PysparseMatrix(matrix=m)

    def __setitem__(self, index, value):
        self.matrix[index] = value.matrix if isinstance(value, PysparseMatrix) else value

    # ---- addition / subtraction --------------------------------------------------------
    def _iadd(self, other, sign=1):
        if self.isSymmetric() and not other.isSymmetric():
            self.matrix.generalize()
        self.matrix.shift(sign, other.getMatrix())
        return self

    def __iadd__(self, other):
        return self._iadd(other)

    def __isub__(self, other):
        return self._iadd(other, sign=-1)

    def __add__(self, other):
        """
        Add two sparse matrices (a new matrix is returned), or add a number to every STORED
        entry of a matrix.

            >>> L = PysparseMatrix(size = 3)
            >>> L.put([3.,10.,numpy.pi,2.5], [0,0,1,2], [2,1,1,0])
            >>> print(L + PysparseIdentityMatrix(size = 3))
             1.000000  10.000000   3.000000
                ---     4.141593      ---
             2.500000      ---     1.000000
            >>> print(L + 3)
                ---    13.000000   6.000000
                ---     6.141593      ---
             5.500000      ---        ---
        """
        if _is_scalar(other):
            if other == 0:
                return self
            L = self.copy()
            val, irow, jcol = L.find()
            L.matrix.update_add_at(other * numpy.ones(val.shape), irow, jcol)
            return L
        if isinstance(other, PysparseMatrix):
            if self.getShape() != other.getShape():
                raise TypeError("Only sparse matrices of the same size may be added")
            L = self.matrix.copy()
            if self.isSymmetric() and not other.isSymmetric():
                L.generalize()
            L.shift(1, other.getMatrix())
            return PysparseMatrix(matrix=L)
        return NotImplemented

    def __sub__(self, other):
        if _is_scalar(other):
            return self.__add__(-other)
        if isinstance(other, PysparseMatrix):
            if self.getShape() != other.getShape():
                raise TypeError("Only sparse matrices of the same size may be subtracted")
            L = self.matrix.copy()
            if self.isSymmetric() and not other.isSymmetric():
                L.generalize()
            L.shift(-1, other.getMatrix())
            return PysparseMatrix(matrix=L)
        return NotImplemented

    # ---- products ------------------------------------------------------------------------
    def __mul__(self, other):
        """
        Multiply a sparse matrix by another sparse matrix, by a vector (the matrix-vector
        product runs on the GPU) or by a scalar.

            >>> L1 = PysparseMatrix(size = 3)
            >>> L1.put([3.,10.,numpy.pi,2.5], [0,0,1,2], [2,1,1,0])
            >>> L2 = PysparseMatrix(size = 3)
            >>> L2.put(numpy.ones(3), numpy.arange(3), numpy.arange(3))
            >>> L2.put([4.38,12357.2,1.1], [2,1,0], [1,0,2])
            >>> tmp = numpy.array(((1.23572000e+05, 2.31400000e+01, 3.00000000e+00),
            ...                    (3.88212887e+04, 3.14159265e+00, 0.00000000e+00),
            ...                    (2.50000000e+00, 0.00000000e+00, 2.75000000e+00)))
            >>> bool(numpy.allclose((L1 * L2).getNumpyArray(), tmp))
            True
        """
        M, N = self.getShape()
        if isinstance(other, PysparseMatrix):
            if N != other.getShape()[0]:
                raise TypeError("Matrices dimensions do not match for product")
            return PysparseMatrix(matrix=spmatrix.matrixmultiply(self.matrix, other.getMatrix()))
        shape = numpy.shape(other)
        if shape == ():
            p = self.matrix.copy()
            p.scale(float(other))
            return PysparseMatrix(matrix=p)
        if shape == (N,):
            x = numpy.ascontiguousarray(other, dtype=numpy.float64)
            y = numpy.empty(M)
            self.matrix.matvec(x, y)
            return y
        raise TypeError("Cannot multiply objects")

    def __imul__(self, other):
        if not _is_scalar(other):
            raise TypeError("In-place multiplication is with scalars only")
        self.matrix.scale(float(other))
        return self

    def __rmul__(self, other):
        # other * A  is really  A^T * other  for a vector
        if isinstance(other, numpy.ndarray) and other.ndim == 1:
            M, N = self.getShape()
            if other.shape != (M,):
                raise TypeError("Cannot multiply objects")
            y = numpy.empty(N)
            self.matrix.matvec_transp(numpy.ascontiguousarray(other, dtype=numpy.float64), y)
            return y
        return self * other

    def matvec(self, x):
        "``A.matvec(x)`` returns ``A * x`` (the one-argument form SciPy's solvers call)."
        return self * x

    # ---- scaling ----------------------------------------------------------------------
    def col_scale(self, v):
        "In-place column scaling: ``A[:,i] *= v[i]``."
        return self.matrix.col_scale(v)

    def row_scale(self, v):
        "In-place row scaling: ``A[i,:] *= v[i]``."
        return self.matrix.row_scale(v)

    # ---- bulk element access ------------------------------------------------------------
    def find(self):
        """
        The matrix in coordinate (triplet) format, three NumPy arrays:

            >>> L = PysparseMatrix(size = 3)
            >>> L.put([3.,10.,numpy.pi,2.5], [0,0,1,2], [2,1,1,0])
            >>> (val,irow,jcol) = L.find()
            >>> irow.tolist(), jcol.tolist()
            ([0, 0, 1, 2], [1, 2, 1, 0])
        """
        return self.matrix.find()

    def _default_ids(self, value, id1, id2):
        nrow, ncol = self.getShape()
        if id2 is None and id1 is not None:
            id2 = id1
        scalar = _is_scalar(value)
        if id1 is None:
            id1 = numpy.arange(nrow if scalar else len(value), dtype=numpy.int64)
        if id2 is None:
            id2 = numpy.arange(ncol if scalar else len(value), dtype=numpy.int64)
        return id1, id2

    def put(self, value, id1=None, id2=None):
        """
        Put elements of ``value`` at positions of the matrix corresponding to ``(id1, id2)``

            >>> L = PysparseMatrix(size = 3)
            >>> L.put( [3.,10.,numpy.pi,2.5], [0,0,1,2], [2,1,1,0] )
            >>> print(L)
                ---    10.000000   3.000000
                ---     3.141593      ---
             2.500000      ---        ---
            >>> L.put(2*numpy.pi, range(3), range(3))
            >>> print(L)
             6.283185  10.000000   3.000000
                ---     6.283185      ---
             2.500000      ---     6.283185

        If ``value`` is a scalar, it acts as the constant vector of appropriate length.  If
        ``id1`` is omitted it is replaced with ``range(nrow)``; if ``id2`` is omitted it is
        ``id1`` when that is present, ``range(ncol)`` otherwise.
        """
        id1, id2 = self._default_ids(value, id1, id2)
        self.matrix.put(value, list(id1) if isinstance(id1, range) else id1,
                        list(id2) if isinstance(id2, range) else id2)

    def putDiagonal(self, vector):
        """
        Put elements of ``vector`` along diagonal of matrix

            >>> L = PysparseMatrix(size = 3)
            >>> L.putDiagonal([3.,10.,numpy.pi])
            >>> L.putDiagonal([10.,3.])
            >>> print(L)
            10.000000      ---        ---
                ---     3.000000      ---
                ---        ---     3.141593
        """
        if _is_scalar(vector):
            ids = numpy.arange(self.getShape()[0])
            self.put(vector, ids, ids)
        else:
            self.matrix.put(vector)

    def take(self, id1=None, id2=None):
        "``val[i] = A[id1[i], id2[i]]`` as a new NumPy array (the diagonal if no indices are given)."
        nrow, ncol = self.getShape()
        if id2 is None and id1 is not None:
            id2 = id1
        if id1 is None:
            id1 = numpy.arange(min(nrow, ncol), dtype=numpy.int64)
        if id2 is None:
            id2 = id1
        vector = numpy.zeros(len(id1), "d")
        self.matrix.take(vector, numpy.asarray(id1), numpy.asarray(id2))
        return vector

    def takeDiagonal(self):
        "The diagonal of the matrix as a NumPy array."
        ids = numpy.arange(min(self.getShape()))
        return self.take(ids, ids)

    def addAt(self, vector, id1, id2):
        """
        Add elements of ``vector`` to the positions in the matrix corresponding to ``(id1,id2)``

            >>> L = PysparseMatrix(size = 3)
            >>> L.put([3.,10.,numpy.pi,2.5], [0,0,1,2], [2,1,1,0])
            >>> L.addAt((1.73,2.2,8.4,3.9,1.23), (1,2,0,0,1), (2,2,0,0,2))
            >>> print(L)
            12.300000  10.000000   3.000000
                ---     3.141593   2.960000
             2.500000      ---     2.200000
        """
        self.matrix.update_add_at(vector, id1, id2)

    def addAtDiagonal(self, vector):
        "Add the components of ``vector`` (or a scalar) to the diagonal elements of the matrix."
        if _is_scalar(vector):
            n = self.getShape()[0]
            ids = numpy.arange(n)
            self.addAt(numpy.full(n, float(vector)), ids, ids)
        else:
            ids = numpy.arange(len(vector))
            self.addAt(vector, ids, ids)

    def getNumpyArray(self):
        "Convert a sparse matrix to a dense NumPy matrix."
        shape = self.getShape()
        indices = numpy.indices(shape)
        return numpy.reshape(self.take(indices[0].ravel(), indices[1].ravel()), shape)

    def exportMmf(self, filename):
        "Exports the matrix to a Matrix Market file of the given filename."
        self.matrix.export_mtx(filename)


class PysparseIdentityMatrix(PysparseMatrix):
    """
    A sparse identity matrix.

        >>> print(PysparseIdentityMatrix(size = 3))
         1.000000      ---        ---
            ---     1.000000      ---
            ---        ---     1.000000
    """

    def __init__(self, size):
        PysparseMatrix.__init__(self, nrow=size, ncol=size, bandwidth=1, symmetric=True)
        ids = numpy.arange(size)
        self.put(numpy.ones(size), ids, ids)


class PysparseSpDiagsMatrix(PysparseMatrix):
    """
    A banded matrix with specified diagonals.

    *Example:* a tridiagonal matrix with 1's on the diagonal, 2's above and -2's below it.

        >>> from numpy import ones
        >>> e = ones(5)
        >>> print(PysparseSpDiagsMatrix(size=5, vals=(-2*e,e,2*e), pos=(-1,0,1)))
         1.000000   2.000000      ---        ---        ---
        -2.000000   1.000000   2.000000      ---        ---
            ---    -2.000000   1.000000   2.000000      ---
            ---        ---    -2.000000   1.000000   2.000000
            ---        ---        ---    -2.000000   1.000000

    Since the ``pos[k]``-th diagonal has ``size-|pos[k]|`` elements, only that many first
    elements of ``vals[k]`` are inserted.  If the matrix is requested to be symmetric,
    elements above the main diagonal are not inserted.
    """

    def __init__(self, size, vals, pos, **kwargs):
        pos = numpy.asarray(pos, dtype=numpy.int64)
        bw = int(max(numpy.abs(pos))) if len(pos) else 0
        diags = size - numpy.abs(pos)
        kwargs.pop("bandwidth", None)
        kwargs.pop("sizeHint", None)
        PysparseMatrix.__init__(self, nrow=size, ncol=size, bandwidth=bw, sizeHint=int(sum(diags)), **kwargs)
        for k in range(len(pos)):
            dk, d = int(diags[k]), int(pos[k])
            v = numpy.asarray(vals[k], dtype=numpy.float64)[:dk]
            if d > 0 and self.isSymmetric():
                continue                       # above the diagonal of a symmetric matrix: not stored
            if d >= 0:
                self.put(v, numpy.arange(dk), d + numpy.arange(dk))
            else:
                self.put(v, -d + numpy.arange(dk), numpy.arange(dk))


class PysparseMatrix4Scipy(PysparseMatrix):
    "A ``PysparseMatrix`` whose ``matvec(x, y)`` has the two-argument ``ll_mat`` signature."

    def matvec(self, x, y):
        return self.matrix.matvec(x, y)
