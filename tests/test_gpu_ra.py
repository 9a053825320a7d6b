"""GPU suite: random access into the device-resident KMC database -- kmx_db_check_kmers /
kmx_db_counters_for_reads against the answers of the reference's CKMCFile::CheckKmer / GetCountersForRead
(tests/golden/ra.json, taken from the compiled reference by make_golden.py) and against the oracle on fresh seeds."""
import hashlib

import numpy as np
import pytest

import cases
import kmcex_b200 as kx
from kmcex_b200 import synth

pytestmark = pytest.mark.gpu


def _oracle_ra(oracle, base, q, reads):
    counts = np.zeros(max(q.size, 1), dtype=np.uint32)
    assert oracle.kmxo_check_kmers(base.encode(), q.ctypes.data, q.size, counts.ctypes.data) == q.size
    flat, off = cases.flat_reads(reads)
    rc = np.zeros(int(off[-1]) + 1, dtype=np.uint32)
    n = oracle.kmxo_counters_for_reads(base.encode(), flat.ctypes.data, off.ctypes.data, len(reads), rc.ctypes.data)
    assert n >= 0
    return counts[: q.size], rc[:n]


@pytest.mark.parametrize("name", sorted(cases.RA_CASES))
def test_gpu_random_access_equals_the_reference(name, ra_dbs, ra_golden, oracle):
    g = ra_golden[name]
    p = g["params"]
    base, sp = ra_dbs(name)
    assert cases.md5_file(base + ".kmc_suf") == g["db_md5"]["kmc_suf"]
    q = cases.ra_queries(sp, p["seed"] + 100)
    reads = cases.ra_reads(sp, p["seed"] + 200)
    db = kx.KmcDatabase(base)
    assert db.info["both_strands"] == (0 if p["one_strand"] else 1) and db.info["signature_len"] == p["sig"]
    counts = db.check_kmers(q)
    per_read = db.counters_for_reads(reads)
    db.close()
    want_c, want_r = _oracle_ra(oracle, base, q, reads)
    assert (counts == want_c).all()
    assert hashlib.md5(counts.tobytes()).hexdigest() == g["check_md5"] and int((counts != 0).sum()) == g["check_hits"]
    assert [a.size for a in per_read] == [max(0, len(r) - sp.k + 1) for r in reads]
    rc = np.concatenate(per_read)
    assert rc.size == g["read_counters"] and (rc == want_r).all()
    assert hashlib.md5(rc.tobytes()).hexdigest() == g["read_counters_md5"]


@pytest.mark.parametrize("k,lut,sig,csz,bins,one_strand", [(31, 7, 6, 2, 3, False), (19, 3, 8, 1, 2, True), (27, 7, 11, 4, 6, False), (15, 3, 10, 2, 1, False),
                                                            (28, 4, 5, 3, 2, False)])
def test_gpu_random_access_on_other_geometries(k, lut, sig, csz, bins, one_strand, oracle, tmp_path):
    base = str(tmp_path / "db")
    cs = 255 if csz == 1 else 1023
    sp = synth.synth_reads_spectrum(30_000, 25, 100, k=k, seed=400 + k, ci=1, cs=cs, device="cpu")
    synth.write_kmc_db(base, sp.kmers, sp.counts, k=k, lut_prefix_length=lut, n_bins=bins, counter_size=csz, min_count=2, max_count=cs - 1,
                       signature_len=sig, signature_bins=True, one_strand=one_strand)
    q = cases.ra_queries(sp, 3000 + k)
    reads = cases.ra_reads(sp, 4000 + k)
    db = kx.KmcDatabase(base)
    counts = db.check_kmers(q)
    rc = np.concatenate(db.counters_for_reads(reads))
    want_c, want_r = _oracle_ra(oracle, base, q, reads)
    assert (counts == want_c).all() and (counts != 0).sum() > 500
    assert rc.size == want_r.size and (rc == want_r).all() and (rc != 0).sum() > 100
    # empty batches, and the calls after a listing / a model build share the resident records
    assert db.check_kmers(np.zeros(0, dtype=np.uint64)).size == 0
    assert db.counters_for_reads([]) == [] and [a.size for a in db.counters_for_reads([b"", b"ACGT"])] == [0, 0]
    kmers, listed = db.list()
    inside = (listed >= 2) & (listed <= cs - 1)
    assert (db.check_kmers(kmers) == np.where(inside, listed, 0)).all()
    db.close()


def test_check_kmers_is_the_exact_count_behind_kmer_to_occ(tmp_path):
    """the use the row names: exact counters next to the model's estimates (tools/accuracy_report.py)"""
    base = str(tmp_path / "db")
    sp = synth.synth_reads_spectrum(200_000, 40, 100, seed=21, ci=1, device="cpu")
    synth.write_kmc_db(base, sp.kmers, sp.counts, lut_prefix_length=7, n_bins=8, min_count=1, signature_bins=True)
    db = kx.KmcDatabase(base)
    exact = db.check_kmers(sp.kmers)
    assert (exact == sp.counts).all()
    m = kx.get_model(1, 1023, 7, 5)
    m.init(db)
    occ = m.kmer_to_occ(sp.kmers)
    same = (occ == exact).mean()
    assert same > 0.5 and (occ == 0).mean() < 0.01               # most answers are exact, hardly any present k-mer reads as 0
    absent = cases.ra_queries(sp, 77)[-3005:-5]
    assert (db.check_kmers(absent) == 0).all()
    m.close()
    db.close()
