"""Accuracy of kmer_to_occ against the exact counters of the KMC database (SURVEY.md 8f row N4 as ground truth).

The reference publishes accuracy claims, not throughput (README.md:3; BASELINE.md section 2 has the survey's own measurement:
78.1 % exact / 21.8 % binned / 0.011 % zero on present k-mers, 0.26 % of absent k-mers answered non-zero).  This script
measures the same rates on the GPU path: the model's answers come from kmx_query_packed, the exact counters from
kmx_db_check_kmers (CKMCFile::CheckKmer on the device-resident database).  Since the model files are byte-identical to
the reference's, these are the reference's rates too.

    python tools/accuracy_report.py [--genome-bp 4600000 --coverage 100 --read-len 101 --ci 2 --seed 1] > profiles/accuracy.json
"""
from __future__ import annotations

import argparse
import json
import os
import sys
import tempfile
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import kmcex_b200 as kx  # noqa: E402
from kmcex_b200 import synth  # noqa: E402


def revcomp(v: np.ndarray, k: int) -> np.ndarray:
    r = np.zeros_like(v)
    t = v.copy()
    for _ in range(k):
        r = (r << np.uint64(2)) | (np.uint64(3) - (t & np.uint64(3)))
        t >>= np.uint64(2)
    return r


def main() -> None:
    ap = argparse.ArgumentParser()
    ap.add_argument("--genome-bp", type=int, default=4_600_000)
    ap.add_argument("--coverage", type=float, default=100)
    ap.add_argument("--read-len", type=int, default=101)
    ap.add_argument("--ci", type=int, default=2)
    ap.add_argument("--seed", type=int, default=1)
    ap.add_argument("--bins", type=int, default=16)
    a = ap.parse_args()
    k = 31
    sp = synth.synth_reads_spectrum(a.genome_bp, a.coverage, a.read_len, k=k, seed=a.seed, ci=a.ci)
    tmp = tempfile.mkdtemp(prefix="kmx_acc_")
    base = os.path.join(tmp, "db")
    synth.write_kmc_db(base, sp.kmers, sp.counts, k=k, lut_prefix_length=7, n_bins=a.bins, min_count=a.ci, signature_bins=True)
    db = kx.KmcDatabase(base)
    m = kx.get_model(a.ci, 1023, 7, 5)
    m.init(db)
    info = m.info
    out_dir = os.path.join(tmp, "model")
    os.makedirs(out_dir)
    m.save(out_dir)
    model_bytes = sum(os.path.getsize(os.path.join(out_dir, f)) for f in ("header", "km.bin", "rest.bin"))

    # present k-mers: every stored k-mer, asked on a random strand
    rng = np.random.default_rng(a.seed + 5)
    flip = rng.integers(0, 2, sp.kmers.size).astype(bool)
    q = np.where(flip, revcomp(sp.kmers, k), sp.kmers)
    t0 = time.time()
    exact = db.check_kmers(sp.kmers).astype(np.int64)          # CheckKmer does not canonicalise: ask the stored form
    t_check = time.time() - t0
    assert (exact == sp.counts).all(), "CheckKmer disagrees with the counters that were written"
    occ = m.kmer_to_occ(q).astype(np.int64)
    same = occ == exact
    zero = occ == 0
    binned = ~same & ~zero
    rel = np.abs(occ[binned] - exact[binned]) / exact[binned]
    by_class = {}
    for name, sel in (("bloom classes (count < ci + bf_num)", exact < a.ci + (1 if a.ci == 1 else 3)), ("coupled arrays / rest", exact >= a.ci + (1 if a.ci == 1 else 3))):
        if sel.any():
            by_class[name] = {"k-mers": int(sel.sum()), "exact": float(same[sel].mean()), "zero": float(zero[sel].mean())}

    # absent k-mers: random 62-bit values and one-base neighbours of stored k-mers, confirmed absent by CheckKmer on both strands
    rnd = rng.integers(0, 1 << 62, 4_000_000, dtype=np.uint64)
    pick = sp.kmers[rng.integers(0, sp.kmers.size, 4_000_000)]
    nb = pick ^ (rng.integers(1, 4, pick.size).astype(np.uint64) << (np.uint64(2) * rng.integers(0, k, pick.size).astype(np.uint64)))
    res = {}
    for name, cand in (("random", rnd), ("one-base neighbours of stored k-mers", nb)):
        canon = np.minimum(cand, revcomp(cand, k))
        truly_absent = db.check_kmers(canon) == 0
        ans = m.kmer_to_occ(cand[truly_absent])
        res[name] = {"queries": int(truly_absent.sum()), "answered_non_zero": float((ans != 0).mean())}

    report = {
        "what": "kmer_to_occ (GPU) against CKMCFile::CheckKmer exact counters (GPU, kmx_db_check_kmers)",
        "database": {"genome_bp": a.genome_bp, "coverage": a.coverage, "read_len": a.read_len, "ci": a.ci, "seed": a.seed, "k": k,
                     "k-mers": int(sp.kmers.size), "bins": a.bins, "binned_by": "KMC signature (len 7)"},
        "model": {"n_hash": 7, "n_bits": 5, "cs": 1023, "bytes": model_bytes, "bytes_per_kmer": model_bytes / sp.kmers.size,
                  "rest_kmers": int(info["rest_kmers"]), "space_vs_8B_kmer_plus_2B_count": 10.0 * sp.kmers.size / model_bytes},
        "present": {"queries": int(q.size), "exact": float(same.mean()), "binned": float(binned.mean()), "zero": float(zero.mean()),
                    "binned_mean_rel_err": float(rel.mean()) if rel.size else 0.0, "by_class": by_class},
        "absent": res,
        "check_kmers_per_s_host_to_host": sp.kmers.size / t_check,
    }
    print(json.dumps(report, indent=1))
    m.close()
    db.close()


if __name__ == "__main__":
    main()
