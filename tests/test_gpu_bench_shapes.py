"""GPU suite: the NAMED BENCH SHAPES (BASELINE.json configs[1] RS-shaped, configs[2] HC14-shaped) built on the GPU from the
seeded generator and compared with the digests tests/golden/make_bench_golden.py took from the UNMODIFIED reference on the
same databases: the parity gate where the numbers are quoted (bench.py repeats it inside every run)."""
import os

import numpy as np
import pytest

import kmcex_b200 as kx
from kmcex_b200 import workloads as wl

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("name", ["rs", "hc14"])
def test_bench_shape_equals_the_reference(name, tmp_path, monkeypatch):
    monkeypatch.setattr(wl, "CACHE", str(tmp_path))
    g = wl.golden_for(name)
    assert g is not None
    meta = wl.ensure_db(name)                               # generated on the GPU here, on the CPU when the golden was taken
    assert meta["db_md5"] == g["db_md5"], "the seeded generator must give the same database on every device"
    assert meta["query_md5"] == g["query_md5"] and meta["n_kmers"] == g["n_kmers"]
    m = kx.get_model(meta["ci"], 1023, 7, 5)
    m.init(meta["db"])
    out = str(tmp_path / "model")
    os.makedirs(out)
    m.save(out)
    assert wl.model_digests(out) == g["model_md5"]
    i = m.info
    assert (i["insert_accepted"], i["rest_kmers"]) == (g["insert_accepted"], g["rest_kmers"])
    # attempts: the reference offers the stale slot-0 item of an empty trailing bucket to the other arrays again
    # (kmodel.hpp:520-540); the kernel knows it is rejected everywhere and appends the duplicate directly
    assert 0 <= g["insert_attempts"] - i["insert_attempts"] <= 5 * 4
    q = np.fromfile(meta["queries"], dtype=np.uint64, count=g["occ_n"])
    occ = m.kmer_to_occ(q)
    assert int((occ != 0).sum()) == g["occ_nonzero"] and wl.occ_digest(occ) == g["occ_md5"]
    m.close()
