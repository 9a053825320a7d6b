#!/usr/bin/env python
"""microseconds per grid-wide barrier on this GPU, for the block shapes the persistent insert kernel can use"""
import ctypes as C
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import kmcex_b200 as kx  # noqa: E402

lib = kx.lib()
kx._lib.check(lib.kmx_set_device(0))
for threads, per_sm in ((256, 4), (512, 2), (1024, 1), (256, 1)):
    for mode, name in ((0, "cg grid.sync"), (1, "counter barrier")):
        us = C.c_float(0)
        kx._lib.check(lib.kmx_microbench_grid_barrier(mode, threads, per_sm, 2000, C.byref(us)))
        print(f"{name:16s} {per_sm * 148:4d} blocks x {threads:4d} threads: {us.value:6.2f} us per barrier")
for n_counters in (1, 4, 16, 64):
    ns = C.c_float(0)
    kx._lib.check(lib.kmx_microbench_hot_atomic(n_counters, 9, C.byref(ns)))
    print(f"returning atomicAdd, 4736 warps x 9 on {n_counters:3d} address(es): {ns.value:6.2f} ns per atomic ({ns.value * 4736 * 9 / 1e3:7.1f} us for the lot)")
