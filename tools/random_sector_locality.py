#!/usr/bin/env python
"""How much of the random-sector cost is address translation / DRAM page locality: the microbenchmark of
tools/random_sector_peaks.py at large footprints, with the accesses of co-running threads confined to a window."""
import ctypes as C
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import kmcex_b200 as kx  # noqa: E402

lib = kx.lib()
kx._lib.check(lib.kmx_set_device(0))
out = {"unit": "G accesses/s (one 32-byte sector each)", "per_item": 7, "results": []}
kinds = {0: "load8", 2: "red_or64"}
for fp_mb in (4096, 16384):
    for win_mb in (0, 2, 32, 512):
        for kind, name in kinds.items():
            n_items = 1 << 25
            ms = C.c_float(0)
            kx._lib.check(lib.kmx_microbench_windowed(kind, fp_mb << 20, win_mb << 20, n_items, 3, C.byref(ms)))
            g = n_items * 7 / (ms.value * 1e-3) / 1e9
            out["results"].append({"kind": name, "footprint_mb": fp_mb, "window_mb": win_mb, "ms": ms.value, "g_accesses_per_s": g})
            print(f"{name:9s} footprint {fp_mb:6d} MiB window {win_mb:4d} MiB: {g:8.1f} G acc/s")
os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
with open(os.path.join(ROOT, "gpurun_out", "random_sector_locality.json"), "w") as f:
    json.dump(out, f, indent=1)
