#!/usr/bin/env python
"""host-side timeline (KMX_TRACE=1) of KModel::init from the files of the bench database, for a few reader-thread counts"""
import os
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
if len(sys.argv) > 1 and sys.argv[1] == "child":
    import kmcex_b200 as kx
    from kmcex_b200 import workloads as bench
    meta = bench.ensure_db(sys.argv[2])
    for i in range(4):
        t0 = time.perf_counter()
        m = kx.get_model(meta["ci"], 1023, 7, 5)
        m.init(meta["db"])
        m.sync()
        print(f"init {i}: {1e3 * (time.perf_counter() - t0):.3f} ms", file=sys.stderr)
        m.close()
else:
    for readers in ("4", "8"):
        print(f"--- KMX_READERS={readers}")
        env = dict(os.environ, KMX_TRACE="1", KMX_READERS=readers)
        r = subprocess.run([sys.executable, __file__, "child", sys.argv[1] if len(sys.argv) > 1 else "rs"], env=env, capture_output=True, text=True)
        print("\n".join(r.stderr.splitlines()[-40:]))
