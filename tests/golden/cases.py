"""The seeded cases behind the golden fixtures: shared by make_golden.py (which runs the
UNMODIFIED reference on them, in the authoring container) and by the tests (which rebuild the
same databases from the seeds and compare against the committed digests)."""
from __future__ import annotations

import hashlib
import os

import numpy as np

# name -> generator parameters.  All use the read simulator (integer arithmetic only), on the CPU
# device, so that the database bytes are identical wherever they are regenerated.
CASES = {
    # one partial batch; ci = 1 (single Bloom filter pair); lut_prefix_length 3, one bin
    "tiny_ci1": dict(genome_bp=20_000, coverage=30, read_len=100, seed=1, ci=1, lut=3, bins=1),
    # ci = 2 (three Bloom filter pairs, probe order 1,0,2); lut_prefix_length 7, four bins
    "small_ci2": dict(genome_bp=200_000, coverage=40, read_len=100, seed=2, ci=2, lut=7, bins=4),
    # more than one batch of n_bits * 2^18 array k-mers, partial last batch with empty trailing
    # buckets (exercises the stale-slot duplicate of kmodel.hpp:520-540)
    "multi_ci1": dict(genome_bp=1_700_000, coverage=12, read_len=100, seed=3, ci=1, lut=3, bins=3),
}
MODEL = dict(cs=1023, n_hash=7, n_bits=5)
N_PRESENT, N_ABSENT, QUERY_SEED = 20000, 20000, 7


def md5_file(path: str) -> str:
    h = hashlib.md5()
    with open(path, "rb") as f:
        for blk in iter(lambda: f.read(1 << 22), b""):
            h.update(blk)
    return h.hexdigest()


def make_case_db(name: str, out_dir: str):
    """(db_base, Spectrum): writes <out_dir>/<name>.kmc_pre/.kmc_suf unless already there"""
    from kmcex_b200 import synth
    p = CASES[name]
    os.makedirs(out_dir, exist_ok=True)
    base = os.path.join(out_dir, name)
    sp = synth.synth_reads_spectrum(p["genome_bp"], p["coverage"], p["read_len"], seed=p["seed"], ci=p["ci"], cs=MODEL["cs"], device="cpu")
    synth.write_kmc_db(base, sp.kmers, sp.counts, k=sp.k, lut_prefix_length=p["lut"], n_bins=p["bins"], min_count=p["ci"],
                       max_count=MODEL["cs"])
    return base, sp


def case_queries(sp) -> np.ndarray:
    from kmcex_b200 import synth
    return synth.neighbour_rich_queries(sp, N_PRESENT, N_ABSENT, seed=QUERY_SEED)


N_DIRTY = 6000


def case_ascii_queries(sp) -> np.ndarray:
    """(N_DIRTY, k) uint8 strings the way reads deliver them: a third clean, a third with one `N`, a third with one
    lower-case base -- the reference never validates a query (tools.hpp:63-76,160-167)"""
    from kmcex_b200 import synth
    q = case_queries(sp)[:N_DIRTY]
    a = synth.to_ascii(q, sp.k).copy()
    rng = np.random.default_rng(QUERY_SEED + 1)
    kind = rng.integers(0, 3, a.shape[0])
    pos = rng.integers(0, sp.k, a.shape[0])
    rows = np.arange(a.shape[0])
    n_rows = rows[kind == 1]
    a[n_rows, pos[n_rows]] = ord("N")
    l_rows = rows[kind == 2]
    a[l_rows, pos[l_rows]] = a[l_rows, pos[l_rows]] + 32          # lower case
    return np.ascontiguousarray(a)
