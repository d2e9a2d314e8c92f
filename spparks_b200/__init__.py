"""spparks_b200 — B200-native colour-sweep rejection-KMC for SPPARKS lattices.

The product is libspparks_cuda (spparks_b200/csrc, C ABI in include/spparks_cuda.h); this package
is the Python host side used by tests and bench.py.  No CPU fallback exists.
"""
from ._lib import SpkCudaError, load  # noqa: F401
from .sweep import LATTICES, MODELS, LatticeSweep  # noqa: F401
