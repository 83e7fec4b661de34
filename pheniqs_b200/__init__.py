"""pheniqs_b200 — B200-native PAMLD / MDD barcode classification behind Pheniqs' decoder interface.

The product is the shared library libpheniqs_b200.so (C ABI in include/pheniqs_b200.h, CUDA
kernels for sm_100a in pheniqs_b200/csrc). This package is the thin host-side layer used by
the tests and the benchmark; importing it never builds or substitutes anything.
"""
from .binding import ConfigurationError, PheniqsError, LIBRARY_PATH  # noqa: F401
from .decoder import DecoderChain, compile_job, load_job, adjust_job, shard_range, all_reduce_accumulators, RESULT_DTYPE, COMPACT_DTYPE  # noqa: F401

__version__ = "0.1.0"
