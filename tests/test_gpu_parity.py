"""GPU suite: the CUDA path behind the C ABI (libkmx.so) against the golden digests taken from
the unmodified reference and against the oracle, bit for bit: KMC listing, header / km.bin /
rest.bin, kmer_to_occ outputs and the per-query path classes."""
import hashlib
import os

import numpy as np
import pytest

import cases
import kmcex_b200 as kx
from kmcex_b200 import synth

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def built(case_dbs, tmp_path_factory):
    """GPU-built and saved model per case (cached)"""
    cache = {}

    def get(name):
        if name not in cache:
            base, sp = case_dbs(name)
            p = cases.CASES[name]
            m = kx.get_model(p["ci"], cases.MODEL["cs"], cases.MODEL["n_hash"], cases.MODEL["n_bits"])
            m.init(base)
            out = str(tmp_path_factory.mktemp(name + "_gpu_model"))
            m.save(out)
            cache[name] = (m, out, sp, base)
        return cache[name]
    return get


def test_extension_is_loaded_and_sees_a_b200():
    lib = kx.lib()
    assert lib.kmx_device_count() >= 1
    loaded = open("/proc/self/maps").read()
    assert "libkmx.so" in loaded


@pytest.mark.parametrize("name", ["tiny_ci1", "small_ci2", "multi_ci1"])
def test_gpu_listing_equals_reference_listing(name, case_dbs, golden, oracle):
    base, sp = case_dbs(name)
    db = kx.KmcDatabase(base)
    kmers, counts = db.list()
    db.close()
    assert kmers.size == sp.kmers.size
    rec = np.zeros(kmers.size, dtype=np.dtype([("k", "<u8"), ("c", "<u4")]))
    rec["k"], rec["c"] = kmers, counts
    assert hashlib.md5(rec.tobytes()).hexdigest() == golden[name]["listing_md5"]
    ok = np.zeros(kmers.size, dtype=np.uint64)
    oc = np.zeros(kmers.size, dtype=np.uint32)
    assert oracle.kmxo_list(base.encode(), ok.ctypes.data, oc.ctypes.data, kmers.size, None, None) == kmers.size
    assert (ok == kmers).all() and (oc == counts).all()


@pytest.mark.parametrize("name", ["tiny_ci1", "small_ci2", "multi_ci1"])
def test_gpu_build_is_byte_identical_to_the_reference(name, built, golden):
    m, out, sp, base = built(name)
    g = golden[name]
    for f in ("header", "km.bin", "rest.bin"):
        assert os.path.getsize(os.path.join(out, f)) == g["model_bytes"][f], f
        assert cases.md5_file(os.path.join(out, f)) == g["model_md5"][f], f
    i = m.info
    assert i["total_kmers"] == sp.kmers.size and i["k"] == 31
    assert i["insert_accepted"] + i["rest_kmers"] >= i["km_kmers"] - (1 << 4)     # every array k-mer lands somewhere
    assert i["insert_attempts"] >= i["insert_accepted"] > 0


@pytest.mark.parametrize("name", ["tiny_ci1", "small_ci2", "multi_ci1"])
def test_gpu_query_equals_reference_outputs(name, built, golden, oracle):
    m, out, sp, base = built(name)
    g = golden[name]
    q = cases.case_queries(sp)
    assert hashlib.md5(q.tobytes()).hexdigest() == g["query_md5"]
    occ = m.kmer_to_occ(q)
    assert occ[:64].tolist() == g["occ_head"]
    assert hashlib.md5(occ.tobytes()).hexdigest() == g["occ_md5"]
    # a model LOADED from the files answers the same (get_model(dir), kmodel.hpp:680-696) and holds the same device arrays
    m2 = kx.get_model(out)
    assert (m2.kmer_to_occ(q) == occ).all()
    assert m2.checksum() == m.checksum() and len(set(m.checksum())) == 4
    # ASCII entry point == packed entry point (the 2-bit encode happens on the device)
    sub = q[:5000]
    assert (m2.kmer_to_occ(synth.to_ascii(sub, 31)) == occ[:5000]).all()
    strs = ["".join(map(chr, row)) for row in synth.to_ascii(sub[:50], 31)]
    assert m2.kmer_to_occ(strs).tolist() == occ[:50].tolist()
    assert m2.kmer_to_occ(strs[0]) == int(occ[0])
    # path classes against the oracle's classification
    h = oracle.kmxo_load(out.encode())
    want = np.zeros(q.size, dtype=np.int32)
    oracle.kmxo_query_path(h, q.ctypes.data, q.size, want.ctypes.data)
    oracle.kmxo_free(h)
    got = m2.query_path(q)
    assert (got == want).all()
    assert set(np.unique(got)) >= {1, 2, 3, 4}            # the fast paths are all exercised
    m2.close()


@pytest.mark.parametrize("name", ["tiny_ci1", "small_ci2", "multi_ci1"])
def test_ascii_queries_with_N_and_lower_case_equal_the_reference(name, built, golden, oracle):
    """the reference never validates a query: a byte that is not C/G/T counts as A for the canonical-form decision and the
    rest lookup, but the filters are probed with hashes of the RAW bytes when the string itself is the canonical form
    (tools.hpp:63-76,160-167; kmodel.hpp:373-390,625-671).  Golden = the reference's own kmer_to_occ(vector<string>)."""
    m, out, sp, base = built(name)
    g = golden[name]
    qa = cases.case_ascii_queries(sp)
    assert hashlib.md5(qa.tobytes()).hexdigest() == g["ascii_query_md5"]
    occ = m.kmer_to_occ(qa)
    assert int((occ != 0).sum()) == g["ascii_occ_nonzero"]
    assert hashlib.md5(occ.tobytes()).hexdigest() == g["ascii_occ_md5"]
    h = oracle.kmxo_load(out.encode())
    want = np.zeros(qa.shape[0], dtype=np.int32)
    oracle.kmxo_query_ascii(h, qa.ctypes.data, qa.shape[1], qa.shape[0], want.ctypes.data)
    oracle.kmxo_free(h)
    assert (occ == want).all()
    # the same strings at a wide stride (the unstaged pack kernel) and as Python strings
    wide = np.full((qa.shape[0], 80), ord("x"), dtype=np.uint8)
    wide[:, :31] = qa
    assert (m.kmer_to_occ(wide) == want).all()
    strs = [bytes(row).decode() for row in qa[:200]]
    assert m.kmer_to_occ(strs).tolist() == want[:200].tolist()


@pytest.mark.parametrize("d", [15_900_007, (1 << 32) + 12_345, 11_000_000_003])
def test_device_addressing_beyond_2_to_32_bits(d, oracle):
    """hash -> exact modulo -> word / bit addressing with array lengths no small database reaches (the NA12878 shape has
    bit_array_length ~ 1.1e10): positions against the oracle's hash % d, and nothing aliases"""
    import ctypes as C
    rng = np.random.default_rng(d & 0xFFFF)
    n, k = 150_000, 31
    kmers = rng.integers(0, 1 << 62, n, dtype=np.uint64)
    seeds = np.array([kx.lib().kmx_host_seed(i) for i in (0, 1, 5, 6, 7, 34, 127)], dtype=np.uint32)
    pos = np.zeros(n * seeds.size, dtype=np.uint64)
    counts = np.zeros(3, dtype=np.uint64)
    kx._lib.check(kx.lib().kmx_selftest_positions(kmers.ctypes.data, n, k, d, seeds.ctypes.data, seeds.size, pos.ctypes.data, counts.ctypes.data))
    pos = pos.reshape(n, seeds.size)
    sample = rng.integers(0, n, 3000)
    for i in sample:
        for j, sd in enumerate(seeds):
            assert int(pos[i, j]) == oracle.kmxo_hash_packed(int(kmers[i]), k, int(sd)) % d
    assert int(pos.max()) < d and (d < (1 << 32) or int(pos.max()) >= (1 << 32))
    distinct = np.unique(pos).size
    assert counts[0] == n * seeds.size           # every position reads back as set (tag bit and filter bit)
    assert counts[1] == distinct and counts[2] == distinct


def test_every_present_kmer_and_all_its_neighbours(built, oracle):
    """all present k-mers + the 8 neighbours of a sample (drives the disambiguation path, kmodel.hpp:286-359)"""
    m, out, sp, base = built("small_ci2")
    k = 31
    mask = np.uint64((1 << 62) - 1)
    sample = sp.kmers[:: max(1, sp.kmers.size // 20000)]
    nb = [((sample << np.uint64(2)) & mask) | np.uint64(b) for b in range(4)]
    nb += [(sample >> np.uint64(2)) | (np.uint64(b) << np.uint64(2 * (k - 1))) for b in range(4)]
    q = np.concatenate([sp.kmers] + nb)
    h = oracle.kmxo_load(out.encode())
    want = np.zeros(q.size, dtype=np.int32)
    oracle.kmxo_query_packed(h, q.ctypes.data, q.size, want.ctypes.data)
    path = np.zeros(q.size, dtype=np.int32)
    oracle.kmxo_query_path(h, q.ctypes.data, q.size, path.ctypes.data)
    oracle.kmxo_free(h)
    got = m.kmer_to_occ(q)
    assert (got == want).all()
    assert (path >= 5).sum() > 0          # neighbour vote / multi-candidate paths were taken


def test_gpu_build_equals_oracle_on_a_fresh_seed(oracle, tmp_path):
    """a database that is in no fixture: ci=2, 7-symbol LUT, 5 bins, random strand queries"""
    base = str(tmp_path / "db")
    sp = synth.synth_reads_spectrum(150_000, 35, 100, seed=1234, ci=2, device="cpu")
    synth.write_kmc_db(base, sp.kmers, sp.counts, lut_prefix_length=7, n_bins=5, min_count=2)
    ora = str(tmp_path / "ora")
    gpu = str(tmp_path / "gpu")
    os.makedirs(ora)
    os.makedirs(gpu)
    stats = np.zeros(3, dtype=np.int64)
    assert oracle.kmxo_build(base.encode(), 2, 1023, 7, 5, ora.encode(), stats.ctypes.data) == 0
    m = kx.get_model(2, 1023, 7, 5)
    m.init_KModel(base)
    m.save_model(gpu)
    for f in ("header", "km.bin", "rest.bin"):
        assert cases.md5_file(os.path.join(gpu, f)) == cases.md5_file(os.path.join(ora, f)), f
    i = m.info
    assert i["insert_attempts"] == stats[0] and i["insert_accepted"] == stats[1] and i["rest_kmers"] == stats[2]


def test_other_geometries_equal_the_oracle(oracle, tmp_path):
    """n_hash / n_bits other than 7 / 5 take the generic kernels"""
    base = str(tmp_path / "db")
    sp = synth.synth_reads_spectrum(40_000, 30, 100, seed=77, ci=1, device="cpu")
    synth.write_kmc_db(base, sp.kmers, sp.counts, lut_prefix_length=3, n_bins=2, min_count=1)
    q = synth.neighbour_rich_queries(sp, 5000, 5000, seed=3)
    for nh, nb in ((6, 4), (8, 3), (7, 1)):
        ora = str(tmp_path / f"ora_{nh}_{nb}")
        gpu = str(tmp_path / f"gpu_{nh}_{nb}")
        os.makedirs(ora)
        os.makedirs(gpu)
        stats = np.zeros(3, dtype=np.int64)
        assert oracle.kmxo_build(base.encode(), 1, 1023, nh, nb, ora.encode(), stats.ctypes.data) == 0
        m = kx.get_model(1, 1023, nh, nb)
        m.init(base)
        m.save(gpu)
        for f in ("header", "km.bin", "rest.bin"):
            assert cases.md5_file(os.path.join(gpu, f)) == cases.md5_file(os.path.join(ora, f)), (nh, nb, f)
        h = oracle.kmxo_load(ora.encode())
        want = np.zeros(q.size, dtype=np.int32)
        oracle.kmxo_query_packed(h, q.ctypes.data, q.size, want.ctypes.data)
        oracle.kmxo_free(h)
        assert (m.kmer_to_occ(q) == want).all(), (nh, nb)
        m.close()


@pytest.mark.parametrize("k,lut,csz,cs,ci,nh,nb", [(27, 3, 2, 1023, 1, 7, 5), (23, 3, 1, 255, 2, 7, 5), (31, 7, 3, 1023, 2, 6, 4),
                                                   (19, 3, 2, 1023, 1, 8, 3), (15, 3, 2, 1023, 1, 7, 2)])
def test_other_k_and_counter_sizes_equal_the_oracle(k, lut, csz, cs, ci, nh, nb, oracle, tmp_path):
    """k != 31, 1/3-byte counters, cs = 255: record layout, rest prefix length and hash tail all change"""
    base = str(tmp_path / "db")
    sp = synth.synth_reads_spectrum(50_000, 30, 100, k=k, seed={27: 11, 23: 12, 31: 13, 19: 14, 15: 16}[k], ci=ci, cs=cs, device="cpu")
    synth.write_kmc_db(base, sp.kmers, sp.counts, k=k, lut_prefix_length=lut, n_bins=3, counter_size=csz, min_count=ci, max_count=cs)
    ora, gpu = str(tmp_path / "ora"), str(tmp_path / "gpu")
    os.makedirs(ora)
    os.makedirs(gpu)
    assert oracle.kmxo_build(base.encode(), ci, cs, nh, nb, ora.encode(), None) == 0
    db = kx.KmcDatabase(base)
    assert db.info["k"] == k and db.info["counter_size"] == csz
    kmers, counts = db.list()
    ok_k = np.zeros(kmers.size, dtype=np.uint64)
    ok_c = np.zeros(kmers.size, dtype=np.uint32)
    assert oracle.kmxo_list(base.encode(), ok_k.ctypes.data, ok_c.ctypes.data, kmers.size, None, None) == kmers.size
    assert (ok_k == kmers).all() and (ok_c == counts).all()
    m = kx.get_model(ci, cs, nh, nb)
    m.init(db)
    m.save(gpu)
    for f in ("header", "km.bin", "rest.bin"):
        assert cases.md5_file(os.path.join(gpu, f)) == cases.md5_file(os.path.join(ora, f)), f
    q = synth.neighbour_rich_queries(sp, 5000, 5000, seed=k)
    h = oracle.kmxo_load(ora.encode())
    want = np.zeros(q.size, dtype=np.int32)
    oracle.kmxo_query_packed(h, q.ctypes.data, q.size, want.ctypes.data)
    oracle.kmxo_free(h)
    assert (m.kmer_to_occ(q) == want).all()
    assert (kx.get_model(gpu).kmer_to_occ(synth.to_ascii(q[:3000], k)) == want[:3000]).all()


def test_empty_and_tiny_batches(built):
    m, out, sp, base = built("tiny_ci1")
    assert m.kmer_to_occ(np.zeros(0, dtype=np.uint64)).size == 0
    assert m.kmer_to_occ([]).size == 0
    one = m.kmer_to_occ(sp.kmers[:1])
    assert one.shape == (1,)
    # a batch larger than one pipeline step of the host path (4 Mi queries), uneven tail
    big = np.resize(sp.kmers, (1 << 22) + 12345)
    occ = m.kmer_to_occ(big)
    assert (occ[: sp.kmers.size] == m.kmer_to_occ(sp.kmers)).all()
    assert (occ[-12345:] == m.kmer_to_occ(big[-12345:])).all()


@pytest.mark.parametrize("env", [{"KMX_TEST_EPOCH_START": "16370"}, {"KMX_RESV_LOG2": "12", "KMX_CLAIM_LOG2": "15"},
                                 {"KMX_CLAIM_FIRST": "1"}, {"KMX_CLAIM_FIRST": "1", "KMX_RESV_LOG2": "12", "KMX_CLAIM_LOG2": "15"},
                                 {"KMX_STREAM_CELLS": "7"}, {"KMX_STREAM_CELLS": "3", "KMX_CLAIM_FIRST": "1"},
                                 {"KMX_MERGED_PASSES": "1"}, {"KMX_MERGED_PASSES": "0"}, {"KMX_MERGED_PASSES": "1", "KMX_TEST_EPOCH_START": "16370"},
                                 {"KMX_MERGED_PASSES": "0", "KMX_TEST_EPOCH_START": "16370"}, {"KMX_MERGED_PASSES": "0", "KMX_RESV_LOG2": "12", "KMX_CLAIM_LOG2": "15"},
                                 {"KMX_MERGED_PASSES": "1", "KMX_RESV_LOG2": "12", "KMX_CLAIM_LOG2": "15"},
                                 {"KMX_MERGED_PASSES": "1", "KMX_CLAIM_FIRST": "1", "KMX_STREAM_CELLS": "3"}])
def test_rare_paths_keep_parity(env, case_dbs, golden, tmp_path, monkeypatch):
    """paths that only large inputs reach: reservation-epoch wrap-around,
    heavily aliased reservation / claim tables (aliasing may only delay decisions, never change them), the
    claim-first phase order, the L2 evict-first probes used for models that do not fit the L2, and both ways of
    resolving contested items (classic two-barrier iterations, merged reserve/commit passes)"""
    for k, v in env.items():
        monkeypatch.setenv(k, v)
    for name in ("small_ci2", "multi_ci1"):
        base, sp = case_dbs(name)
        m = kx.get_model(cases.CASES[name]["ci"], cases.MODEL["cs"], cases.MODEL["n_hash"], cases.MODEL["n_bits"])
        m.init(base)
        out = str(tmp_path / (name + "_".join(env)))
        os.makedirs(out)
        m.save(out)
        for f in ("header", "km.bin", "rest.bin"):
            assert cases.md5_file(os.path.join(out, f)) == golden[name]["model_md5"][f], (env, name, f)
        m.close()


def test_reference_error_corners_are_reported_not_computed(tmp_path):
    # fewer than 8 k-mers in a Bloom class: the reference aborts in new uint8_t[0]{0} (kmodel.hpp:413-417)
    base = str(tmp_path / "db")
    kmers = np.arange(1000, 1100, dtype=np.uint64) * np.uint64(0x9E3779B97F4A7C15) & np.uint64((1 << 62) - 1)
    kmers = np.unique(kmers)
    counts = np.full(kmers.size, 50, dtype=np.uint32)
    counts[:3] = 1
    synth.write_kmc_db(base, kmers, counts, lut_prefix_length=3, min_count=1)
    m = kx.get_model(1, 1023, 7, 5)
    with pytest.raises(kx.KmxError) as e:
        m.init(base)
    assert e.value.code == 6
    with pytest.raises(kx.KmxError):
        m.save(str(tmp_path))             # not initialised
