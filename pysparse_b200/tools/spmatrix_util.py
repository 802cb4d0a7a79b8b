"""Small helpers around the matrix types (reference: ``pysparse/tools/spmatrix_util.py``;
the VTK viewers are not provided)."""
import random

from ..sparse import spmatrix

__all__ = ["bytesToString", "printInfo", "ll_mat_rand"]


def bytesToString(n):
    "Human-readable size, as the reference prints it ('%d Bytes', '%.1f Kbytes', ...)."
    if n < 1024:
        return "%d Bytes" % n
    for unit in ("Kbytes", "Mbytes"):
        n /= 1024.0
        if n < 1024:
            return "%.1f %s" % (n, unit)
    return "%.1f Gbytes" % (n / 1024.0)


def storage_bytes(mat):
    "Bytes of matrix storage by the reference's own formulas (spmatrix_util.py:24-30), or None."
    t = type(mat)
    if t is spmatrix.LLMatType:
        return 4 * mat.shape[0] + 16 * mat.nnz
    if t is spmatrix.SSSMatType:
        return 12 * mat.shape[0] + 12 * (mat.nnz - mat.shape[0])
    if t is spmatrix.CSRMatType:
        return 4 * mat.shape[0] + 12 * mat.nnz
    return None


def printInfo(mat, name):
    "Print the statistics block of the reference (name, type, dimensions, #non-zeros, storage)."
    t = type(mat)
    if t is spmatrix.LLMatType:
        typeName = "LL symmetric" if mat.issym else "LL general"
    elif t is spmatrix.SSSMatType:
        typeName = "SSS"
    elif t is spmatrix.CSRMatType:
        typeName = "CSR"
    else:
        typeName = "Unknown"
    nbytes = storage_bytes(mat)
    print("Matrix statistics:")
    print("%-20s: %s" % ("name", name))
    print("%-20s: %s" % ("type", typeName))
    print("%-20s: %dx%d" % ("dimensions", mat.shape[0], mat.shape[1]))
    print("%-20s: %d" % ("#non-zeros", mat.nnz))
    print("%-20s: %s" % ("storage", bytesToString(nbytes) if nbytes is not None else "Unknown"))
    print()


def ll_mat_rand(n, m, density):
    """A general n-by-m ``ll_mat`` with at most ``n*m*density`` random entries in [0, 1)
    (positions drawn with replacement, like the reference)."""
    nnz = int(density * n * m)
    A = spmatrix.ll_mat(n, m, max(nnz, 1))
    for _ in range(nnz):
        A[random.randrange(n), random.randrange(m)] = random.random()
    return A
