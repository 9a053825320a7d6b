#!/usr/bin/env python
"""A/B builds of the insert kernel (block size, grid barrier); load one with KMX_LIB_PATH=kmcex_b200/<name>"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from kmcex_b200 import build as b  # noqa: E402

VARIANTS = {
    "libkmx_gb0.so": ["KMX_GRIDBAR=0"],
    "libkmx_t512.so": ["KMX_INS_THREADS=512"],
    "libkmx_t1024.so": ["KMX_INS_THREADS=1024"],
}
for name, defs in VARIANTS.items():
    if sys.argv[1:] and name not in sys.argv[1:]:
        continue
    print(b.build_lib(force=True, out=os.path.join(ROOT, "kmcex_b200", name), defines=defs))
