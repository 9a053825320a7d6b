#!/usr/bin/env python
"""read-only streaming rate of this GPU (the ceiling of count_kernel) next to count_kernel's own rate on the HC14-shaped database"""
import ctypes as C
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import kmcex_b200 as kx  # noqa: E402

lib = kx.lib()
kx._lib.check(lib.kmx_set_device(0))
for bps in (2, 4, 8):
    ms = C.c_float(0)
    kx._lib.check(lib.kmx_microbench_stream_read(2 << 30, bps, 10, C.byref(ms)))
    print(f"read-only stream, {bps} blocks/SM: {(2 << 30) / (ms.value * 1e-3) / 1e9:8.0f} GB/s")
