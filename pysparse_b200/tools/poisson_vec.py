"""Vectorised Poisson assembly (reference: ``pysparse/tools/poisson_vec.py``): the same matrices
as ``tools.poisson``, built with ``put`` on index arrays instead of element by element."""
from .poisson import (poisson1d_sym_vec, poisson1d_vec, poisson2d_sym_blk_vec, poisson2d_sym_vec,  # noqa: F401
                      poisson2d_vec, poisson2d_vec2)
