#!/usr/bin/env python
"""Measure the random-sector roofline denominators on this GPU and write profiles/random_sector_peaks.json."""
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
kinds = {0: "load8", 1: "red_or32", 2: "red_or64"}
for fp_mb in (16, 64, 512, 4096):
    for kind, name in kinds.items():
        n_items = 1 << 24
        ms = C.c_float(0)
        kx._lib.check(lib.kmx_microbench_random(kind, fp_mb << 20, n_items, 5, C.byref(ms)))
        g = n_items * 7 / (ms.value * 1e-3) / 1e9
        out["results"].append({"kind": name, "footprint_mb": fp_mb, "ms": ms.value, "g_accesses_per_s": g, "gb_per_s_32B": g * 32})
        print(f"{name:9s} footprint {fp_mb:5d} MiB: {g:8.1f} G acc/s  = {g * 32:8.0f} GB/s of 32-byte sectors")
os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
with open(os.path.join(ROOT, "gpurun_out", "random_sector_peaks.json"), "w") as f:
    json.dump(out, f, indent=1)
