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
