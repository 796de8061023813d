"""walt_b200 -- B200-native WALT mapping engine (CUDA kernels behind a C ABI).

The product is `walt_b200/lib/libwaltb200.so` (built from `walt_b200/csrc/`) and the C++
host program `walt_b200/bin/walt`.  This Python package is a thin ctypes binding of the C ABI
(`include/walt_b200.h`) used by the tests and by bench.py; it contains no mapping logic and
has no CPU fallback: importing works anywhere, every mapping call needs the CUDA library and
a device.
"""
from .engine import (BEST_DT, CAND_DT, PAIR_DT, Engine, Group, WaltError, lib_path, load_library,  # noqa: F401
                     pack_reads, CT00, CT01, GA10, GA11)
