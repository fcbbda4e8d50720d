"""B200-native sparse polynomial multiplier: the hot path of piranha's
``series_multiplier<polynomial<Cf, kronecker_monomial<>>>`` as hand-written sm_100a CUDA kernels behind a C ABI.

Layout:
  csrc/   CUDA kernels + the C ABI implementation (built into lib/libpiranha_b200.so by ./build.sh)
  host/   C++ host mirror of the reference interface for this path (kronecker_array, polynomial, series_multiplier)
  abi.py  ctypes binding of include/piranha_b200.h used by tests/ and bench.py

There is no CPU implementation in this package; without the CUDA library and a B200 every call raises.
"""
from .abi import (ALGO_NAMES, PK_ALGO_AUTO, PK_ALGO_DENSE, PK_ALGO_HASH, PK_ALGO_ZONE, PK_ALGO_BLOCK, Context, DevicePoly, PkError,
                  digest_terms, load_library)

__all__ = ["Context", "DevicePoly", "PkError", "load_library", "digest_terms", "PK_ALGO_AUTO", "PK_ALGO_DENSE",
           "PK_ALGO_HASH", "PK_ALGO_ZONE", "PK_ALGO_BLOCK", "ALGO_NAMES"]
