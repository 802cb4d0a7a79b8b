"""pysparse_b200.sparse -- drop-in for ``pysparse.sparse`` (module ``spmatrix``)."""
from . import spmatrix  # noqa: F401  (C extension; ImportError if the native build is missing)
