"""pysparse_b200.eigen -- drop-in for ``pysparse.eigen``: the ``jdsym`` module."""
from . import jdsym  # noqa: F401
