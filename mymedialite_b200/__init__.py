"""mymedialite_b200 -- B200-native matrix-factorization engine behind MyMediaLite's recommender API.

The compute path is libmmlb200.so (hand-written sm_100a CUDA behind a C ABI, include/mmlb200.h). This
package is the host-side mirror of the reference's interface used by the tests and the benchmark; it never
falls back to a CPU implementation."""
from . import _capi  # noqa: F401
from ._capi import MmlError, MFParams  # noqa: F401
