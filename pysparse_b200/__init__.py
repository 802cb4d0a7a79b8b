"""pysparse_b200 -- B200-native (sm_100a) implementation of PySparse's SpMV / Krylov hot path."""
__version__ = "0.1.0"
