#!/usr/bin/env python
"""build the bench model once, then answer the bench query set a few times (device-resident): the command ncu profiles
for the retrieval kernels"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
from kmcex_b200 import workloads as bench  # noqa: E402
import kmcex_b200 as kx  # noqa: E402

w = sys.argv[1] if len(sys.argv) > 1 else "rs"
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 3
meta = bench.ensure_db(w)
m = kx.get_model(meta["ci"], 1023, 7, 5)
m.init(meta["db"])
q = np.fromfile(meta["queries"], dtype=np.uint64)[: 1 << 24]
dev = torch.device("cuda", 0)
qd = torch.from_numpy(q.astype(np.int64)).to(dev)
out = torch.empty(q.size, dtype=torch.int32, device=dev)
st = torch.cuda.Stream()
ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
for r in range(reps):
    ev[0].record(st)
    m.query_device(qd.data_ptr(), q.size, out.data_ptr(), st.cuda_stream)
    ev[1].record(st)
    torch.cuda.synchronize()
    print(f"{w}: {q.size / ev[0].elapsed_time(ev[1]) / 1e6:.3f} G queries/s")
