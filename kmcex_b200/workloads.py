"""The named benchmark shapes (BASELINE.json configs) as seeded synthetic KMC databases: shared by bench.py,
tests/golden/make_bench_golden.py (which pins them against the unmodified reference) and the GPU tests.
Support code, not the hot path."""
from __future__ import annotations

import hashlib
import json
import os

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

WORKLOADS = {
    # name: (synth shape, ci, lut_prefix_length, bins, BASELINE.json config it stands for)
    "small": ("small", 2, 7, 4, "test-sized (200 kbp, 40x)"),
    "cfg1": ("cfg1", 1, 3, 8, "configs[0]: 1M-read 100bp, ci1"),
    "rs": ("rs", 2, 7, 16, "configs[1]: GAGE-RS-shaped synthetic (4.6 Mbp, 100x, 101bp) k31 nh7 nb5 ci2"),
    "hc14": ("hc14", 1, 7, 64, "configs[2]: GAGE-HC14-shaped synthetic (88 Mbp, 40x) k31 nh7 nb5 ci1"),
    "wgs350": ("wgs350", 2, 7, 128, "scale check towards configs[3]: 350 Mbp synthetic genome, 30x, k31 nh7 nb5 ci2"),
    "na12878": ("na12878", 2, 7, 512, "configs[3]: NA12878-shaped synthetic (3.1 Gbp, 30x, 101bp) k31 nh7 nb5 ci2"),
}
# shapes generated bin group by bin group on the GPU (kmcex_b200.synth.make_db_streamed): (genome_bp, coverage, read_len)
STREAMED = {"na12878": (3_100_000_000, 30, 101)}
SWEEP_POOL = 100_000_000          # distinct queries behind the configs[4] sweep
CACHE = os.environ.get("KMX_BENCH_CACHE", "/tmp/kmx_bench")
MODEL_FILES = ("header", "km.bin", "rest.bin")
GOLDEN = os.path.join(ROOT, "tests", "golden", "bench_shapes.json")


def md5_file(path: str) -> str:
    h = hashlib.md5()
    with open(path, "rb") as f:
        for blk in iter(lambda: f.read(1 << 24), b""):
            h.update(blk)
    return h.hexdigest()


def _write_meta(meta_path: str, meta: dict) -> None:
    tmp = meta_path + f".{os.getpid()}"
    with open(tmp, "w") as f:
        json.dump(meta, f)
    os.replace(tmp, meta_path)


def ensure_db(workload: str, seed: int = 1) -> dict:
    """generate (or reuse) the synthetic KMC database + a query set for it; returns paths, sizes and md5 digests"""
    from . import synth
    shape, ci, lut, bins, _ = WORKLOADS[workload]
    d = os.path.join(CACHE, f"{workload}_s{seed}")
    base = os.path.join(d, "db")
    meta_path = os.path.join(d, "meta.json")
    if os.path.exists(meta_path):
        with open(meta_path) as f:
            meta = json.load(f)
        if "db_md5" in meta:
            return meta
    os.makedirs(d, exist_ok=True)
    if workload in STREAMED:
        import shutil
        import torch
        g, cov, rl = STREAMED[workload]
        need = int(g * 1.4 * 8) + (8 << 30)
        if shutil.disk_usage(d).free < need:
            raise SystemExit(f"{workload}: {need >> 30} GiB of scratch space needed under {CACHE} (set KMX_BENCH_CACHE)")
        r = synth.make_db_streamed(base, g, cov, rl, seed=seed, ci=ci, lut_prefix_length=lut, n_bins=bins, n_present=SWEEP_POOL // 2)
        torch.cuda.empty_cache()
        q = synth.mixed_queries(r["present"], SWEEP_POOL, seed=seed + 100)
        n_kmers = int(r["n_kmers"])
    else:
        sp = synth.make_db(base, shape, seed=seed, ci=ci, lut_prefix_length=lut, n_bins=bins)
        # query set: 50 % present (random strand) / 50 % absent + neighbours, BASELINE.json configs[4] mix
        n_q = 1 << 24
        q = synth.neighbour_rich_queries(sp, n_q // 2, n_q // 2 - (n_q // 2) // 4, seed=seed + 100)
        n_kmers = int(sp.kmers.size)
    q.tofile(os.path.join(d, "queries.u64"))
    meta = {"db": base, "queries": os.path.join(d, "queries.u64"), "n_kmers": n_kmers, "n_queries": int(q.size), "ci": ci,
            "suffix_bytes": os.path.getsize(base + ".kmc_suf"), "prefix_bytes": os.path.getsize(base + ".kmc_pre"),
            "db_md5": {"kmc_pre": md5_file(base + ".kmc_pre"), "kmc_suf": md5_file(base + ".kmc_suf")},
            "query_md5": md5_file(os.path.join(d, "queries.u64")), "seed": seed}
    _write_meta(meta_path, meta)
    return meta


def golden_for(workload: str, seed: int = 1):
    """digests of the reference's model for this shape (tests/golden/bench_shapes.json, written by
    tests/golden/make_bench_golden.py from the UNMODIFIED reference), or None when the shape is not pinned"""
    if not os.path.exists(GOLDEN):
        return None
    with open(GOLDEN) as f:
        return json.load(f).get(f"{workload}_s{seed}")


def model_digests(model_dir: str) -> dict:
    return {f: md5_file(os.path.join(model_dir, f)) for f in MODEL_FILES}


def occ_digest(occ: np.ndarray) -> str:
    return hashlib.md5(np.ascontiguousarray(occ, dtype=np.int32).tobytes()).hexdigest()
