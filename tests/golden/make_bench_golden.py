"""Pin the NAMED BENCH SHAPES (BASELINE.json configs[1] RS-shaped, configs[2] HC14-shaped) against the UNMODIFIED
reference: regenerate the seeded databases of kmcex_b200/workloads.py, run oracle/_ref/ref_driver (compiled from
/root/reference by oracle/Makefile) on them and record the digests of the database, of header / km.bin / rest.bin and
of the kmer_to_occ answers in tests/golden/bench_shapes.json.  Run in the authoring container only:

    python tests/golden/make_bench_golden.py [rs hc14]

bench.py and tests/test_gpu_bench_shapes.py compare the GPU build of the same seeded database with these digests
(the generator is integer / exactly-rounded arithmetic only, so the database bytes are the same on CPU and CUDA)."""
from __future__ import annotations

import ctypes as C
import json
import os
import subprocess
import sys
import time

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
from kmcex_b200 import workloads as wl  # noqa: E402

REF = os.path.join(ROOT, "oracle", "_ref", "ref_driver")
OCC_N = 1 << 22          # queries of the bench query set whose answers are pinned


def main() -> None:
    if not os.path.exists(REF):
        raise SystemExit("oracle/_ref/ref_driver missing: run `make -C oracle ref` (needs /root/reference)")
    names = sys.argv[1:] or ["rs", "hc14"]
    out = {}
    if os.path.exists(wl.GOLDEN):
        with open(wl.GOLDEN) as f:
            out = json.load(f)
    for name in names:
        seed = 1
        t0 = time.time()
        meta = wl.ensure_db(name, seed)
        t1 = time.time()
        ci = meta["ci"]
        mdir = os.path.join(wl.CACHE, f"{name}_s{seed}", "ref_model")
        os.makedirs(mdir, exist_ok=True)
        r = subprocess.run([REF, "build", meta["db"], mdir, str(ci), "1023", "7", "5"], capture_output=True, text=True, check=True)
        t2 = time.time()
        qf = os.path.join(wl.CACHE, f"{name}_s{seed}", "golden_q.u64")
        q = np.fromfile(meta["queries"], dtype=np.uint64, count=OCC_N)
        q.tofile(qf)
        of = qf + ".occ"
        subprocess.run([REF, "query", mdir, qf, "31", of, str(os.cpu_count() or 1)], capture_output=True, text=True, check=True)
        occ = np.fromfile(of, dtype=np.int32)
        entry = {
            "workload": wl.WORKLOADS[name][4], "seed": seed, "ci": ci, "n_kmers": meta["n_kmers"],
            "db_md5": meta["db_md5"], "query_md5": meta["query_md5"],
            "model_md5": wl.model_digests(mdir),
            "model_bytes": {f: os.path.getsize(os.path.join(mdir, f)) for f in wl.MODEL_FILES},
            "occ_n": int(q.size), "occ_md5": wl.occ_digest(occ), "occ_nonzero": int((occ != 0).sum()), "occ_sum": int(occ.astype(np.int64).sum()),
            "reference_stdout_tail": r.stdout.strip().splitlines()[-1],
        }
        # the CPU restatement (oracle/libkmx_oracle.so) on the same database: pinned on the bench shapes as well
        ora = C.CDLL(os.path.join(ROOT, "oracle", "libkmx_oracle.so"))
        ora.kmxo_build.argtypes = [C.c_char_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_char_p, C.c_void_p]
        odir = os.path.join(wl.CACHE, f"{name}_s{seed}", "oracle_model")
        os.makedirs(odir, exist_ok=True)
        stats = np.zeros(3, dtype=np.int64)
        assert ora.kmxo_build(meta["db"].encode(), ci, 1023, 7, 5, odir.encode(), stats.ctypes.data) == 0
        entry["oracle_equals_reference"] = wl.model_digests(odir) == entry["model_md5"]
        entry["insert_attempts"], entry["insert_accepted"], entry["rest_kmers"] = (int(x) for x in stats)
        assert entry["oracle_equals_reference"], f"{name}: the oracle's files differ from the reference's"
        out[f"{name}_s{seed}"] = entry
        print(name, f"generate {t1 - t0:.0f}s, reference build {t2 - t1:.0f}s", json.dumps(entry)[:300])
        with open(wl.GOLDEN, "w") as f:
            json.dump(out, f, indent=1, sort_keys=True)


if __name__ == "__main__":
    main()
