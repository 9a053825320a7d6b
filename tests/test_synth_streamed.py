"""The bin-group-by-bin-group database generator behind the NA12878-shaped workload (BASELINE.json configs[3]):
on a small genome, on the CPU, the files it writes must list -- through the oracle's restatement of
CKMCFile::ReadNextKmer (kmc_file.cpp:428-515) -- exactly the unique canonical k-mers it says it wrote."""
import ctypes as C

import numpy as np

from kmcex_b200 import synth


def test_streamed_database_lists_what_was_written(oracle, tmp_path):
    base = str(tmp_path / "db")
    r = synth.make_db_streamed(base, 50_000, 30, 101, ci=2, lut_prefix_length=3, n_bins=8, n_groups=4, chunk=1 << 13,
                               n_present=500, device="cpu")
    n = r["n_kmers"]
    km = np.zeros(n + 8, dtype=np.uint64)
    ct = np.zeros(n + 8, dtype=np.uint32)
    k, total = C.c_int32(), C.c_uint64()
    got = oracle.kmxo_list(base.encode(), km.ctypes.data, ct.ctypes.data, n + 8, C.byref(k), C.byref(total))
    assert got == n == total.value and k.value == 31
    km, ct = km[:n], ct[:n]
    assert np.unique(km).size == n                                   # merged duplicates: every record is a different k-mer
    assert ct.min() >= 2 and ct.max() <= 1023                        # -ci2, saturation at cs
    assert all(np.bincount(ct)[c] >= 8 for c in (2, 3, 4))           # every Bloom class is populated (kmodel.hpp:413-417)
    t = synth.torch.from_numpy(km.astype(np.int64))
    assert bool((synth.canonical_packed(t, 31) == t).all())          # canonical form, as KMC stores it
    assert np.isin(r["present"], km).all()
    # same seed, other chunking / grouping: the same database
    base2 = str(tmp_path / "db2")
    r2 = synth.make_db_streamed(base2, 50_000, 30, 101, ci=2, lut_prefix_length=3, n_bins=8, n_groups=2, chunk=1 << 15, device="cpu")
    assert r2["n_kmers"] == n
    for ext in (".kmc_pre", ".kmc_suf"):
        assert open(base + ext, "rb").read() == open(base2 + ext, "rb").read()


def test_mixed_queries_have_the_configs4_mix():
    present = np.arange(1000, 3000, dtype=np.uint64) * np.uint64(0x9E3779B97F4A7C15) & np.uint64((1 << 62) - 1)
    q = synth.mixed_queries(present, 4096, seed=3)
    assert q.size == 4096 and q.dtype == np.uint64 and int(q.max()) < (1 << 62)
    t = synth.torch.from_numpy(present.astype(np.int64))
    both = np.concatenate([present, synth.revcomp_packed(t, 31).numpy().astype(np.uint64)])
    frac = np.isin(q, both).mean()
    assert 0.49 <= frac <= 0.52                                       # half of the batch is present k-mers (either strand)
