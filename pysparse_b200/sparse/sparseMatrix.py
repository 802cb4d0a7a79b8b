"""``pysparse.sparse.sparseMatrix``: the abstract base class of the matrix wrappers."""
from .pysparseMatrix import SparseMatrix  # noqa: F401

__all__ = ["SparseMatrix"]
