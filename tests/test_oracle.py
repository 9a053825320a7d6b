"""CPU suite, part 2: the oracle (oracle/kmx_oracle.cpp, a CPU restatement used only as a
checker) reproduces the golden digests that tests/golden/make_golden.py took from the UNMODIFIED
reference, and -- when the compiled reference is present -- its live output."""
import hashlib
import os
import subprocess

import numpy as np
import pytest

import cases

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = os.path.join(ROOT, "oracle", "_ref", "ref_driver")


def _oracle_build(oracle, base, p, out_dir):
    os.makedirs(out_dir, exist_ok=True)
    stats = np.zeros(3, dtype=np.int64)
    rc = oracle.kmxo_build(base.encode(), p["ci"], cases.MODEL["cs"], cases.MODEL["n_hash"], cases.MODEL["n_bits"], out_dir.encode(),
                           stats.ctypes.data)
    assert rc == 0
    return stats


@pytest.mark.parametrize("name", ["tiny_ci1", "small_ci2", "multi_ci1"])
def test_oracle_matches_reference_goldens(name, oracle, case_dbs, golden, tmp_path):
    g = golden[name]
    base, sp = case_dbs(name)
    assert cases.md5_file(base + ".kmc_pre") == g["db_md5"]["kmc_pre"]       # the seeded database is reproducible
    assert cases.md5_file(base + ".kmc_suf") == g["db_md5"]["kmc_suf"]
    # listing order and values (CKMCFile::ReadNextKmer)
    n = sp.kmers.size
    kmers = np.zeros(n, dtype=np.uint64)
    counts = np.zeros(n, dtype=np.uint32)
    got = oracle.kmxo_list(base.encode(), kmers.ctypes.data, counts.ctypes.data, n, None, None)
    assert got == n
    rec = np.zeros(n, dtype=np.dtype([("k", "<u8"), ("c", "<u4")]))
    rec["k"], rec["c"] = kmers, counts
    assert hashlib.md5(rec.tobytes()).hexdigest() == g["listing_md5"]
    # build
    out = str(tmp_path / "model")
    stats = _oracle_build(oracle, base, cases.CASES[name], out)
    for f in ("header", "km.bin", "rest.bin"):
        assert os.path.getsize(os.path.join(out, f)) == g["model_bytes"][f]
        assert cases.md5_file(os.path.join(out, f)) == g["model_md5"][f], f
    assert stats[0] >= stats[1] > 0
    # query
    q = cases.case_queries(sp)
    assert hashlib.md5(q.tobytes()).hexdigest() == g["query_md5"]
    h = oracle.kmxo_load(out.encode())
    assert h
    occ = np.zeros(q.size, dtype=np.int32)
    oracle.kmxo_query_packed(h, q.ctypes.data, q.size, occ.ctypes.data)
    oracle.kmxo_free(h)
    assert occ[:64].tolist() == g["occ_head"]
    assert int((occ != 0).sum()) == g["occ_nonzero"]
    assert hashlib.md5(occ.tobytes()).hexdigest() == g["occ_md5"]
    # strings with N / lower case: the reference hashes their raw bytes when the forward orientation is canonical
    qa = cases.case_ascii_queries(sp)
    assert hashlib.md5(qa.tobytes()).hexdigest() == g["ascii_query_md5"]
    h = oracle.kmxo_load(out.encode())
    occ_a = np.zeros(qa.shape[0], dtype=np.int32)
    oracle.kmxo_query_ascii(h, qa.ctypes.data, qa.shape[1], qa.shape[0], occ_a.ctypes.data)
    oracle.kmxo_free(h)
    assert int((occ_a != 0).sum()) == g["ascii_occ_nonzero"]
    assert hashlib.md5(occ_a.tobytes()).hexdigest() == g["ascii_occ_md5"]
    assert int((occ_a != occ[:occ_a.size]).sum()) == g["ascii_differs_from_clean"] > 0      # reading N as A would be wrong


@pytest.mark.skipif(not os.path.exists(REF), reason="compiled reference (oracle/_ref) not present")
@pytest.mark.parametrize("k,lut,csz,cs,ci,nh,nb", [(31, 7, 2, 1023, 2, 7, 5), (27, 3, 2, 1023, 1, 7, 5), (23, 3, 1, 255, 2, 7, 5),
                                                   (19, 3, 3, 1023, 1, 8, 3), (15, 3, 2, 1023, 1, 6, 2)])
def test_oracle_matches_live_reference_on_a_fresh_seed(k, lut, csz, cs, ci, nh, nb, oracle, tmp_path):
    from kmcex_b200 import synth
    base = str(tmp_path / "db")
    sp = synth.synth_reads_spectrum(60_000, 25, 100, k=k, seed=99 + k, ci=ci, cs=cs, device="cpu")
    synth.write_kmc_db(base, sp.kmers, sp.counts, k=k, lut_prefix_length=lut, n_bins=2, counter_size=csz, min_count=ci, max_count=cs)
    ref_dir, ora_dir = str(tmp_path / "ref"), str(tmp_path / "ora")
    os.makedirs(ref_dir)
    os.makedirs(ora_dir)
    subprocess.run([REF, "build", base, ref_dir, str(ci), str(cs), str(nh), str(nb)], check=True, capture_output=True)
    assert oracle.kmxo_build(base.encode(), ci, cs, nh, nb, ora_dir.encode(), None) == 0
    for f in ("header", "km.bin", "rest.bin"):
        assert cases.md5_file(os.path.join(ref_dir, f)) == cases.md5_file(os.path.join(ora_dir, f)), f
    q = synth.neighbour_rich_queries(sp, 5000, 5000, seed=5)
    qf, of = str(tmp_path / "q.bin"), str(tmp_path / "o.bin")
    q.tofile(qf)
    subprocess.run([REF, "query", ref_dir, qf, str(k), of, "2"], check=True, capture_output=True)
    ref = np.fromfile(of, dtype=np.int32)
    h = oracle.kmxo_load(ref_dir.encode())
    occ = np.zeros(q.size, dtype=np.int32)
    oracle.kmxo_query_packed(h, q.ctypes.data, q.size, occ.ctypes.data)
    oracle.kmxo_free(h)
    assert (occ == ref).all()


# ---- random access: CKMCFile::CheckKmer / GetCountersForRead ---------------------------------------
def _oracle_ra(oracle, base, q, reads):
    counts = np.zeros(q.size, dtype=np.uint32)
    assert oracle.kmxo_check_kmers(base.encode(), q.ctypes.data, q.size, counts.ctypes.data) == q.size
    flat, off = cases.flat_reads(reads)
    k_total = int(off[-1]) + 1
    rc = np.zeros(k_total, dtype=np.uint32)
    n = oracle.kmxo_counters_for_reads(base.encode(), flat.ctypes.data, off.ctypes.data, len(reads), rc.ctypes.data)
    assert n >= 0
    return counts, rc[:n]


@pytest.mark.parametrize("name", sorted(cases.RA_CASES))
def test_oracle_random_access_matches_reference_goldens(name, oracle, ra_dbs, ra_golden):
    g = ra_golden[name]
    base, sp = ra_dbs(name)
    assert cases.md5_file(base + ".kmc_pre") == g["db_md5"]["kmc_pre"] and cases.md5_file(base + ".kmc_suf") == g["db_md5"]["kmc_suf"]
    q = cases.ra_queries(sp, g["params"]["seed"] + 100)
    reads = cases.ra_reads(sp, g["params"]["seed"] + 200)
    assert hashlib.md5(q.tobytes()).hexdigest() == g["query_md5"] and hashlib.md5(b"\n".join(reads)).hexdigest() == g["reads_md5"]
    counts, rc = _oracle_ra(oracle, base, q, reads)
    assert int((counts != 0).sum()) == g["check_hits"] and counts[:32].tolist() == g["check_head"]
    assert hashlib.md5(counts.tobytes()).hexdigest() == g["check_md5"]
    assert rc.size == g["read_counters"] and int((rc != 0).sum()) == g["read_counters_nonzero"]
    assert hashlib.md5(rc.tobytes()).hexdigest() == g["read_counters_md5"]
    if g["params"]["signature_bins"] and g["params"]["min_count"] == 1:
        # ground truth: with a real signature map every stored k-mer is found with its counter
        assert (counts[:6000] == sp.counts[np.searchsorted(sp.kmers, q[:6000])]).all()


@pytest.mark.skipif(not os.path.exists(REF), reason="compiled reference (oracle/_ref) not present")
@pytest.mark.parametrize("k,lut,sig,csz,bins,one_strand", [(31, 7, 6, 2, 3, False), (19, 3, 8, 1, 2, True), (27, 7, 11, 4, 6, False), (15, 3, 10, 2, 1, False)])
def test_oracle_random_access_matches_live_reference(k, lut, sig, csz, bins, one_strand, oracle, tmp_path):
    from kmcex_b200 import synth
    base = str(tmp_path / "db")
    cs = 255 if csz == 1 else 1023
    sp = synth.synth_reads_spectrum(30_000, 25, 100, k=k, seed=300 + k, ci=1, cs=cs, device="cpu")
    synth.write_kmc_db(base, sp.kmers, sp.counts, k=k, lut_prefix_length=lut, n_bins=bins, counter_size=csz, min_count=2, max_count=cs - 1,
                       signature_len=sig, signature_bins=True, one_strand=one_strand)
    q = cases.ra_queries(sp, 1000 + k)
    reads = cases.ra_reads(sp, 2000 + k)
    qf, cf, rf, of = (str(tmp_path / n) for n in ("q.bin", "c.bin", "reads.txt", "rc.bin"))
    q.tofile(qf)
    with open(rf, "wb") as f:
        f.write(b"\n".join(reads) + b"\n")
    subprocess.run([REF, "check", base, qf, cf], check=True, capture_output=True)
    subprocess.run([REF, "reads", base, rf, of], check=True, capture_output=True)
    counts, rc = _oracle_ra(oracle, base, q, reads)
    assert (counts == np.fromfile(cf, dtype=np.uint32)).all()
    assert rc.size == os.path.getsize(of) // 4 and (rc == np.fromfile(of, dtype=np.uint32)).all()
