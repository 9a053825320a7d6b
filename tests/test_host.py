"""CPU suite, part 1: the C ABI loads and exports what include/kmx.h declares; the host-side
arithmetic of the path (hash, canonical form, OccuBin, size formulas, exact modulo, survivor
permutation) agrees with the reference's known answers (tests/golden/kat.txt, printed by the
reference's own functions) and with the oracle."""
import ctypes as C
import os
import random
import re

import numpy as np
import pytest

import kmcex_b200 as kx
from kmcex_b200 import _lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exports_every_declared_symbol():
    text = open(os.path.join(ROOT, "include", "kmx.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    declared = set(re.findall(r"\b(kmx_[a-z0-9_]+)\s*\(", text))
    assert len(declared) >= 25
    lib = kx.lib()
    for name in sorted(declared):
        assert hasattr(lib, name), f"libkmx.so does not export {name}"
    assert declared == set(_lib.SIGNATURES), "ctypes table and kmx.h disagree"


def test_no_torch_types_in_the_abi():
    text = open(os.path.join(ROOT, "include", "kmx.h")).read()
    assert "torch" not in text and "at::" not in text and "std::" not in text


def test_murmur_and_canonical_known_answers(kat_lines, oracle):
    lib = kx.lib()
    code = {"A": 0, "C": 1, "G": 2, "T": 3}
    n_m = n_c = 0
    for parts in kat_lines:
        if parts[0] == "murmur":
            s, seed, want = parts[1], int(parts[2]), int(parts[3])
            assert lib.kmx_host_murmur64(s.encode(), len(s), seed) == want
            assert oracle.kmxo_murmur64(s.encode(), len(s), seed) == want
            v = 0
            for ch in s:
                v = (v << 2) | code[ch]
            assert lib.kmx_host_hash_packed(v, len(s), seed) == want      # the kernels' block/tail split
            n_m += 1
        elif parts[0] == "minkmer":
            s, want = parts[1], parts[2]
            v = w = 0
            for ch in s:
                v = (v << 2) | code[ch]
            for ch in want:
                w = (w << 2) | code[ch]
            assert lib.kmx_host_canonical(v, len(s)) == w
            assert oracle.kmxo_canonical(v, len(s)) == w
            n_c += 1
    assert n_m >= 500 and n_c >= 100
    # the survey's probe values
    assert lib.kmx_host_murmur64(b"ACGTACGTACGTACGTACGTACGTACGTACG", 31, 46757) == 13042456722222963169
    assert lib.kmx_host_murmur64(b"ACGTACGTACGTACGTACGTACGTACGTACG", 31, 46769) == 17098917893835065645
    assert lib.kmx_host_murmur64(b"CGTACGTACGTACGTACGTACGTACGTAC", 29, 46757) == 2953946542005570904


def test_kmc_signature_known_answers(kat_lines, oracle):
    """CKmerAPI::get_signature (kmer_api.h:653-673, mmer.h:33-88): library, oracle and the test-data writer against the reference"""
    from kmcex_b200 import synth
    lib = kx.lib()
    code = {"A": 0, "C": 1, "G": 2, "T": 3}
    n = 0
    for parts in kat_lines:
        if parts[0] != "signature":
            continue
        s, sl, want = parts[1], int(parts[2]), int(parts[3])
        v = 0
        for ch in s:
            v = (v << 2) | code[ch]
        assert lib.kmx_host_signature(v, len(s), sl) == want, (s, sl)
        assert oracle.kmxo_signature(v, len(s), sl) == want, (s, sl)
        assert int(synth.kmc_signatures(np.array([v], dtype=np.uint64), len(s), sl)[0]) == want, (s, sl)
        n += 1
    assert n >= 400


def test_seed_table(oracle):
    lib = kx.lib()
    seeds = [lib.kmx_host_seed(i) for i in range(128)]
    assert seeds[0] == 46757 and seeds[1] == 46769 and seeds[127] == 48163
    assert seeds == [oracle.kmxo_seed(i) for i in range(128)]


def test_occubin_known_answers(kat_lines, oracle):
    lib = kx.lib()
    tables = {}
    for parts in kat_lines:
        if parts[0] != "occubin":
            continue
        mc, nh, occ, b, mean = map(int, parts[1:])
        if (mc, nh) not in tables:
            o2b = np.zeros(mc, dtype=np.int32)
            b2m = np.zeros(1 << nh, dtype=np.int32)
            assert lib.kmx_host_occubin(mc, nh, o2b.ctypes.data, b2m.ctypes.data) == 0
            o2b_o = np.zeros(mc, dtype=np.int32)
            b2m_o = np.zeros(1 << nh, dtype=np.int32)
            assert oracle.kmxo_occubin(mc, nh, o2b_o.ctypes.data, b2m_o.ctypes.data) == 0
            assert (o2b == o2b_o).all() and (b2m == b2m_o).all()
            tables[(mc, nh)] = (o2b, b2m)
        o2b, b2m = tables[(mc, nh)]
        assert o2b[occ] == b
        assert (b if b < (1 << nh) // 4 else b2m[b]) == mean
    assert set(tables) == {(1024, 7), (256, 7), (1024, 6), (65536, 8)}
    # occu_bin.hpp:38-44 writes out of bounds when max_counter < 2^(H-2) + 3 * 2^(H-1): refused
    tmp = np.zeros(4096, dtype=np.int32)
    assert lib.kmx_host_occubin(200, 7, tmp.ctypes.data, tmp.ctypes.data) != 0


def test_size_formulas():
    lib = kx.lib()
    rng = random.Random(5)
    for _ in range(2000):
        counts = [rng.randrange(0, 1 << rng.randrange(1, 40)) for _ in range(3)]
        km = rng.randrange(0, 1 << rng.randrange(1, 40))
        nh = rng.randrange(3, 12)
        bf_num = rng.choice([1, 3])
        arr = (C.c_uint64 * 3)(*counts)
        out = (C.c_uint64 * 8)()
        lib.kmx_host_sizes(arr, bf_num, km, nh, out)
        for i in range(bf_num):
            assert out[i] == int(counts[i] / 5.5 * (nh - 1))           # kmodel.hpp:411, double arithmetic
            assert out[3 + i] == (counts[i] >> 3) * (nh - 2)            # kmodel.hpp:415
        assert out[6] == (km >> 4) * nh and out[7] == (km >> 4) * (nh - 2)   # kmodel.hpp:437-439


def test_fastmod_is_exact():
    lib = kx.lib()
    rng = random.Random(11)
    ds = [1, 2, 3, 7, 8, 63, 64, 65, (1 << 32) - 1, 1 << 32, (1 << 32) + 1, (1 << 61) - 1, 1 << 61, 8400000000, 2863311531]
    ds += [rng.randrange(1, 1 << rng.randrange(1, 62)) for _ in range(300)]
    hs = [0, 1, (1 << 64) - 1, (1 << 63), (1 << 63) - 1] + [rng.randrange(0, 1 << 64) for _ in range(200)]
    for d in ds:
        for h in hs + [d - 1, d, d + 1, 2 * d, 3 * d - 1, ((1 << 64) - 1) // d * d, ((1 << 64) - 1) // d * d - 1]:
            h &= (1 << 64) - 1
            assert lib.kmx_host_fastmod(h, d) == h % d, (h, d)


def test_reorder_closed_form_equals_two_pointer_loop(oracle):
    lib = kx.lib()
    rng = np.random.default_rng(3)
    for trial in range(3000):
        n = int(rng.integers(0, 70)) if trial < 2500 else int(rng.integers(1000, 5000))
        p = rng.choice([0.0, 0.1, 0.5, 0.9, 1.0])
        failed = (rng.random(n) < p).astype(np.uint8)
        a = np.full(n + 1, -1, dtype=np.int32)
        b = np.full(n + 1, -1, dtype=np.int32)
        fa = np.ascontiguousarray(failed) if n else np.zeros(1, np.uint8)
        ra = lib.kmx_host_reorder(fa.ctypes.data, n, a.ctypes.data)
        rb = oracle.kmxo_reorder(fa.ctypes.data, n, b.ctypes.data)
        assert ra == int(failed.sum())
        if n > 0:
            assert ra == rb
            assert (a[:ra] == b[:rb]).all()


def test_python_api_mirrors_the_reference_names():
    for name in ("init", "init_KModel", "save", "save_model", "kmer_to_occ", "show_header_info", "show_kmodel_info"):
        assert hasattr(kx.KModel, name)
    m = kx.get_model(2, 1023, 7, 5)
    i = m.info
    assert (i["ci"], i["cs"], i["n_hash"], i["n_bits"], i["bf_num"]) == (2, 1023, 7, 5, 3)    # kmodel.hpp:50
    assert kx.get_model(1).info["bf_num"] == 1
    with pytest.raises(kx.KmxError):
        kx.get_model(1, 100, 7, 5)        # OccuBin out of bounds in the reference
    with pytest.raises(kx.KmxError):
        kx.get_model("/nonexistent/model/dir")


def test_kmc_header_parse_without_gpu(case_dbs, golden):
    base, sp = case_dbs("tiny_ci1")
    db = kx.KmcDatabase(base)
    i = db.info
    assert i["k"] == 31 and i["lut_prefix_length"] == 3 and i["counter_size"] == 2 and i["kmc_version"] == 0x200
    assert i["total_kmers"] == sp.kmers.size == golden["tiny_ci1"]["n_kmers"]
    assert i["record_bytes"] == 9 and i["suffix_bytes"] == 9 * sp.kmers.size and i["lut_entries"] == 64
    db.close()
    with pytest.raises(kx.KmxError):
        kx.KmcDatabase("/nonexistent/db")


def test_compute_entry_points_fail_loudly_without_gpu(case_dbs):
    if kx.lib().kmx_device_count() > 0:
        pytest.skip("a CUDA device is present")
    base, sp = case_dbs("tiny_ci1")
    m = kx.get_model(1, 1023, 7, 5)
    with pytest.raises(kx.KmxError) as e:
        m.init(base)
    assert e.value.code == 4          # KMX_ENOGPU: no CPU fallback
    with pytest.raises(kx.KmxError):
        m.kmer_to_occ("ACGTACGTACGTACGTACGTACGTACGTACG")
    db = kx.KmcDatabase(base)
    with pytest.raises(kx.KmxError) as e:
        db.check_kmers(sp.kmers[:10])          # random access: no host-side search either
    assert e.value.code == 4
    with pytest.raises(kx.KmxError):
        db.counters_for_reads([b"A" * 40])
    db.close()


def test_counter_range_of_a_database_can_only_be_narrowed(ra_dbs):
    """CKMCFile::SetMinCount / SetMaxCount / ResetMinMaxCounts (kmc_file.cpp:670-734)"""
    base, _ = ra_dbs("ra_k31_range")           # header range [3, 40]
    db = kx.KmcDatabase(base)
    lib = kx.lib()
    assert (db.info["min_count"], db.info["max_count"], db.info["both_strands"], db.info["signature_len"]) == (3, 40, 1, 7)
    assert lib.kmx_db_set_count_range(db._h, 5, 30) == 0 and (db.info["min_count"], db.info["max_count"]) == (5, 30)
    for lo, hi in ((2, 30), (5, 41), (31, 30)):
        assert lib.kmx_db_set_count_range(db._h, lo, hi) != 0
    assert (db.info["min_count"], db.info["max_count"]) == (5, 30)
    assert lib.kmx_db_reset_count_range(db._h) == 0 and (db.info["min_count"], db.info["max_count"]) == (3, 40)
    db.close()


def test_signature_is_strand_symmetric_and_bounded():
    """norm(m) = norm(revcomp(m)) (mmer.h:63-88), so a k-mer and its reverse complement share the signature, hence the bin"""
    lib = kx.lib()
    rng = np.random.default_rng(5)
    for k, sl in ((31, 7), (27, 9), (32, 11), (15, 5), (23, 6)):
        for v in rng.integers(0, 1 << 62, 300, dtype=np.uint64):
            v = int(v) & ((1 << (2 * k)) - 1)
            rc, t = 0, v
            for _ in range(k):
                rc = (rc << 2) | (3 - (t & 3))
                t >>= 2
            a = lib.kmx_host_signature(v, k, sl)
            assert a == lib.kmx_host_signature(rc, k, sl) and a <= 4 ** sl
    assert lib.kmx_host_signature(0, 31, 7) == 4 ** 7            # poly-A: no allowed m-mer
    assert lib.kmx_host_signature(0, 31, 4) == 0xFFFFFFFF        # signature lengths outside 5..11 are refused


def test_team_item_routing_is_a_bijection_onto_the_owners_shards():
    """kmx_host_route: stream position -> (owner of the round-0 array, index in the owner's shard).  Bucket i of every batch
    goes to rank i % n_active (array-owner decomposition, kmodel.hpp:561-565); an owner keeps its buckets back to back."""
    lib = kx.lib()
    K = 1 << 18
    for n_bits, n_active in ((5, 1), (5, 2), (5, 3), (5, 5), (8, 8), (3, 2), (1, 1)):
        seen = {}
        n_buckets = 3 * n_bits + 2                      # three full batches and a partial one
        per_owner = [0] * n_active
        for B in range(n_buckets):
            i = B % n_bits
            want_owner = i % n_active
            for c in (0, 1, K - 1):
                o, at = C.c_int32(-1), C.c_uint64(0)
                lib.kmx_host_route(B * K + c, n_active, n_bits, C.byref(o), C.byref(at))
                assert o.value == want_owner
                assert at.value == per_owner[want_owner] * K + c, (n_bits, n_active, B, c)
                assert (o.value, at.value) not in seen
                seen[(o.value, at.value)] = B * K + c
            per_owner[want_owner] += 1
        # single GPU: the shard is the stream
        o, at = C.c_int32(-1), C.c_uint64(0)
        lib.kmx_host_route(123456789, 1, n_bits, C.byref(o), C.byref(at))
        assert (o.value, at.value) == (0, 123456789)


def test_team_prefix_cuts_partition_the_rest_table():
    """sharded rest build: contiguous prefix ranges, balanced on the histogram, identical on every rank"""
    lib = kx.lib()
    rng = np.random.default_rng(5)
    for trial in range(20):
        map_size = int(rng.choice([64, 1024, 16384]))
        hist = rng.integers(0, 50, map_size).astype(np.uint32)
        if trial % 4 == 0:
            hist[: map_size // 2] = 0                    # empty prefixes at the front
        if trial % 5 == 0:
            hist[:] = 0
            hist[7] = 1000                               # everything in one group
        for world in (1, 2, 3, 8):
            cut = np.zeros(world + 1, dtype=np.uint32)
            off = np.zeros(world + 1, dtype=np.uint64)
            assert lib.kmx_host_prefix_cuts(hist.ctypes.data, map_size, world, cut.ctypes.data, off.ctypes.data) == 0
            assert cut[0] == 0 and cut[world] == map_size and (np.diff(cut.astype(np.int64)) >= 0).all()
            csum = np.concatenate([[0], np.cumsum(hist.astype(np.uint64))])
            assert (off == csum[cut]).all()              # a rank's run starts where the prefixes before its range end
            total = int(csum[-1])
            sizes = np.diff(off.astype(np.int64))
            assert sizes.sum() == total
            if total and hist.max() * world < total:     # no giant group: shares stay within one group of the ideal
                assert sizes.max() <= total // world + int(hist.max()) + 1


def test_bench_shapes_are_pinned_by_the_reference():
    """tests/golden/bench_shapes.json: digests of the unmodified reference's model for the named bench shapes, with the
    oracle checked against them when they were taken"""
    from kmcex_b200 import workloads as wl
    for name in ("rs", "hc14"):
        g = wl.golden_for(name)
        assert g is not None and g["oracle_equals_reference"] is True
        assert set(g["model_md5"]) == set(wl.MODEL_FILES) and g["occ_n"] == 1 << 22 and len(g["db_md5"]["kmc_suf"]) == 32
        assert g["insert_attempts"] >= g["insert_accepted"] > 0 and g["workload"] == wl.WORKLOADS[name][4]


def test_device_selection_fails_loudly_without_a_gpu():
    """kmx_set_devices validates its list against the devices present; on a box without a GPU every ordinal is refused
    (KMX_ENOGPU), an empty list (= single-GPU behaviour) is accepted"""
    lib = kx.lib()
    n = lib.kmx_device_count()
    assert lib.kmx_set_devices(None, 0) == 0
    bad = (C.c_int * 2)(n, n + 1)                       # ordinals beyond what is present
    assert lib.kmx_set_devices(bad, 2) == 4             # KMX_ENOGPU
    assert lib.kmx_set_devices(bad, 9) == 1             # KMX_EARG: more than 8 devices
    if n >= 1:
        twice = (C.c_int * 2)(0, 0)
        assert lib.kmx_set_devices(twice, 2) == 1       # the same device listed twice
    assert lib.kmx_team_steps() == 4 and lib.kmx_team_blob_bytes() == 256
    assert lib.kmx_set_devices(None, 0) == 0
