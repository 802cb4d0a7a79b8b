"""pysparse_b200.precon -- drop-in for ``pysparse.precon`` (module ``precon``)."""
from . import precon  # noqa: F401
