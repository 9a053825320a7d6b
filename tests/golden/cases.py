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


# ---- random access (CKMCFile::CheckKmer / GetCountersForRead) -------------------------------------
# databases binned by KMC signature with a real signature map (what a KMC run writes); "hash_bins" reuses the layout of the
# cases above (bins by a hash of the value, all-zero signature map): CheckKmer then only ever searches bin 0
RA_CASES = {
    "ra_k31_sig7": dict(genome_bp=60_000, coverage=25, read_len=100, seed=5, k=31, lut=3, bins=5, sig=7, counter_size=2, cs=1023,
                        min_count=1, max_count=1023, signature_bins=True, one_strand=False),
    "ra_k27_sig9": dict(genome_bp=40_000, coverage=30, read_len=90, seed=6, k=27, lut=3, bins=3, sig=9, counter_size=1, cs=255,
                        min_count=1, max_count=255, signature_bins=True, one_strand=False),
    # forward-strand database: GetCountersForRead does not canonicalise; 3-byte counters
    "ra_k23_one_strand": dict(genome_bp=30_000, coverage=20, read_len=80, seed=7, k=23, lut=7, bins=2, sig=5, counter_size=3, cs=1023,
                              min_count=1, max_count=1023, signature_bins=True, one_strand=True),
    # the header's counter range is narrower than the records: CheckKmer reports those k-mers as absent (kmc_file.cpp:1426-1434)
    "ra_k31_range": dict(genome_bp=30_000, coverage=20, read_len=100, seed=8, k=31, lut=3, bins=1, sig=7, counter_size=2, cs=1023,
                         min_count=3, max_count=40, signature_bins=True, one_strand=False),
    "ra_k31_hash_bins": dict(genome_bp=30_000, coverage=20, read_len=100, seed=9, k=31, lut=7, bins=4, sig=7, counter_size=2, cs=1023,
                             min_count=1, max_count=1023, signature_bins=False, one_strand=False),
}


def make_ra_db(name: str, out_dir: str):
    from kmcex_b200 import synth
    p = RA_CASES[name]
    os.makedirs(out_dir, exist_ok=True)
    base = os.path.join(out_dir, name)
    sp = synth.synth_reads_spectrum(p["genome_bp"], p["coverage"], p["read_len"], k=p["k"], seed=p["seed"], ci=1, cs=p["cs"], device="cpu")
    synth.write_kmc_db(base, sp.kmers, sp.counts, k=p["k"], lut_prefix_length=p["lut"], n_bins=p["bins"], counter_size=p["counter_size"],
                       min_count=p["min_count"], max_count=p["max_count"], signature_len=p["sig"], signature_bins=p["signature_bins"],
                       one_strand=p["one_strand"])
    return base, sp


def ra_queries(sp, seed: int) -> np.ndarray:
    """packed k-mers for CheckKmer: stored k-mers, their reverse complements, one-base neighbours, random values"""
    rng = np.random.default_rng(seed)
    k = sp.k
    mask = np.uint64((1 << (2 * k)) - 1)
    present = sp.kmers[rng.integers(0, sp.kmers.size, 6000)]
    rc = np.zeros_like(present[:2000])
    t = present[:2000].copy()
    for _ in range(k):
        rc = (rc << np.uint64(2)) | (np.uint64(3) - (t & np.uint64(3)))
        t >>= np.uint64(2)
    nb = present[2000:4000] ^ (rng.integers(1, 4, 2000).astype(np.uint64) << (np.uint64(2) * rng.integers(0, k, 2000).astype(np.uint64)))
    rnd = rng.integers(0, 1 << 62, 3000, dtype=np.uint64) & mask
    edge = np.array([0, int(mask), int(sp.kmers[0]), int(sp.kmers[-1]), int(sp.kmers[-1]) + 1 & int(mask)], dtype=np.uint64)
    return np.ascontiguousarray(np.concatenate([present, rc, nb, rnd, edge]))


def ra_reads(sp, seed: int) -> list[bytes]:
    """reads for GetCountersForRead: genome segments (either strand), with N, other non-ACGT bytes and lower case sprinkled in,
    plus reads shorter than k, of exactly k, and an empty one"""
    rng = np.random.default_rng(seed)
    k = sp.k
    g = sp.genome
    letters = np.frombuffer(b"ACGT", dtype=np.uint8)
    comp = np.frombuffer(b"TGCA", dtype=np.uint8)
    reads = []
    for i in range(120):
        n = int(rng.integers(k - 6, 3 * k + 20))
        at = int(rng.integers(0, g.size - n))
        seg = g[at: at + n]
        r = comp[seg[::-1]].copy() if i % 2 else letters[seg].copy()
        mode = i % 6
        if mode == 1 and n > 3:
            r[rng.integers(0, n, int(rng.integers(1, 4)))] = ord("N")
        elif mode == 2:
            sel = rng.integers(0, n, n // 3)
            r[sel] = r[sel] + 32                                  # lower case is a base (kmer_api.h:270-273)
        elif mode == 3 and n > 3:
            r[rng.integers(0, n, 2)] = np.frombuffer(b".R", dtype=np.uint8)
        elif mode == 4 and n > 2 * k:
            r[k - 1] = ord("n")                                    # the first k-mer and the k - 1 after it are invalid
            r[n - 1] = ord("N")
        reads.append(r.tobytes())
    reads += [b"", letters[g[100: 100 + k]].tobytes(), letters[g[200: 200 + k - 1]].tobytes(), b"N" * (k + 5), b"A" * (2 * k), b"ACA" * k, b"T" * (k + 3)]
    return reads


def flat_reads(reads: list[bytes]):
    """(uint8 bases back to back, int64 offsets[n + 1]) -- the layout of kmx_db_counters_for_reads"""
    off = np.zeros(len(reads) + 1, dtype=np.int64)
    np.cumsum([len(r) for r in reads], out=off[1:])
    return np.frombuffer(b"".join(reads) + b"\0", dtype=np.uint8), off
