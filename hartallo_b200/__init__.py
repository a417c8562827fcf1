"""hartallo_b200 -- B200 (sm_100a) implementation of hartallo's H.264 encoder pixel hot path.

The product is ``libhl_b200.so`` (hand-written CUDA behind the C-ABI of ``include/hlb200.h``).  This Python package is
only the binding used by tests and bench.py; it has NO CPU fallback: importing ``hartallo_b200.lib`` without the built
library, or calling it without a CUDA device, raises.
"""
from .lib import Hlb200Error, load  # noqa: F401
