#!/usr/bin/env python
"""random 32-byte-sector loads / reductions issued by GPU 0 into GPU 1's memory over NVLink, next to the local rates:
the cost of owner-routed bits (north_star's address-range sharding of the Bloom filters) against replicate-then-OR"""
import ctypes as C
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import kmcex_b200 as kx  # noqa: E402

lib = kx.lib()
if lib.kmx_device_count() < 2:
    raise SystemExit("needs 2 GPUs")
kx._lib.check(lib.kmx_set_device(0))
kinds = {0: "load8", 1: "red_or32", 2: "red_or64"}
n_items = 1 << 22
for fp_mb in (64, 1024):
    for kind, name in kinds.items():
        ms_l, ms_r = C.c_float(0), C.c_float(0)
        kx._lib.check(lib.kmx_microbench_random(kind, fp_mb << 20, n_items, 3, C.byref(ms_l)))
        kx._lib.check(lib.kmx_microbench_peer_random(kind, 0, 1, fp_mb << 20, n_items, 3, C.byref(ms_r)))
        g = lambda ms: n_items * 7 / (ms.value * 1e-3) / 1e9
        print(f"{name:9s} footprint {fp_mb:5d} MiB: local {g(ms_l):7.1f} G acc/s   remote (GPU0 -> GPU1 over NVLink) {g(ms_r):7.2f} G acc/s")
