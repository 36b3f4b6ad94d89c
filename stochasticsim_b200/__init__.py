"""stochasticsim_b200 -- host-side Python mirror of the C ABI in include/ssb200.h.

The product is libssb200.so (hand-written sm_100a CUDA behind `extern "C"`) plus the two drop-in
C mains in stochasticsim_b200/host/.  This package is only the ctypes binding used by tests and
bench.py; it contains no computation and NO fallback: if the shared library has not been built
(`make`, or `__graft_entry__.build()`) importing any entry point raises.
"""
from ._lib import lib, SSBError, Context, TncCarry, LIB_PATH, check  # noqa: F401
from . import tnc  # noqa: F401
from . import spike  # noqa: F401

__all__ = ["lib", "SSBError", "Context", "TncCarry", "LIB_PATH", "check", "tnc", "spike"]
