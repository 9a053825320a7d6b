"""GPU suite: the k-mer counting stage (FASTQ -> KMC database) that replaces the reference's call of the external
`kmc` binary (main.cpp:136-140).  Checked against a numpy count of the same reads, read back through the
reference's listing semantics (oracle), and fed into the model build."""
import os

import numpy as np
import pytest

import kmcex_b200 as kx
from kmcex_b200 import counter, synth

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("k,ci,cs,crlf", [(31, 1, 1023, False), (31, 2, 1023, True), (27, 1, 255, False), (19, 3, 1023, False)])
def test_counter_equals_numpy_count(k, ci, cs, crlf, oracle, tmp_path):
    fq = str(tmp_path / "reads.fastq")
    want_k, want_c, n_reads = synth.synth_fastq(fq, genome_bp=30_000, coverage=25, read_len=100, k=k, seed=k + ci, crlf=crlf)
    base = str(tmp_path / "db")
    info = counter.count_fastq(fq, base, k=k, ci=ci, cs=cs)
    assert info["n_reads"] == n_reads
    assert info["n_unique"] == want_k.size                                   # (windows with an N are not k-mers)
    keep = want_c >= ci
    want_k, want_c = want_k[keep], np.minimum(want_c[keep], cs)
    assert info["n_kept"] == want_k.size
    # read the database back with the reference's listing semantics
    n = want_k.size
    got_k = np.zeros(n, dtype=np.uint64)
    got_c = np.zeros(n, dtype=np.uint32)
    kk = np.zeros(1, dtype=np.int32)
    assert oracle.kmxo_list(base.encode(), got_k.ctypes.data, got_c.ctypes.data, n, kk.ctypes.data, None) == n
    assert kk[0] == k
    assert (got_k == want_k).all() and (got_c == want_c).all()
    # and with the GPU listing
    db = kx.KmcDatabase(base)
    lk, lc = db.list()
    db.close()
    assert (lk == want_k).all() and (lc == want_c).all()


def test_gzip_input_and_file_lists_count_the_same(tmp_path):
    """the reference's usage text promises "FASTQ format (gziped or not)" (main.cpp:40-41): zlib on the host reads both"""
    import gzip
    import shutil
    fq = str(tmp_path / "reads.fastq")
    want_k, want_c, n_reads = synth.synth_fastq(fq, genome_bp=20_000, coverage=15, read_len=100, k=31, seed=77)
    with open(fq, "rb") as src, gzip.open(fq + ".gz", "wb") as dst:
        shutil.copyfileobj(src, dst)
    a = counter.count_fastq(fq, str(tmp_path / "plain"), k=31, ci=2, cs=1023)
    b = counter.count_fastq(fq + ".gz", str(tmp_path / "gz"), k=31, ci=2, cs=1023)
    assert a == b and a["n_reads"] == n_reads and a["n_kept"] == int((want_c >= 2).sum())
    for ext in (".kmc_pre", ".kmc_suf"):
        assert open(str(tmp_path / "plain") + ext, "rb").read() == open(str(tmp_path / "gz") + ext, "rb").read()
    # two files in one call = their concatenation: every count doubles
    c = counter.count_fastq([fq, fq + ".gz"], str(tmp_path / "both"), k=31, ci=2, cs=1023)
    assert c["n_reads"] == 2 * n_reads and c["n_kept"] == int((2 * want_c >= 2).sum())


def test_fastq_to_model_end_to_end(oracle, tmp_path):
    """BASELINE.json configs[0] in miniature: FASTQ -> count -> model build -> query, on the GPU"""
    fq = str(tmp_path / "reads.fastq")
    want_k, want_c, _ = synth.synth_fastq(fq, genome_bp=60_000, coverage=30, read_len=100, k=31, seed=9)
    base = str(tmp_path / "db")
    counter.count_fastq([fq], base, k=31, ci=1, cs=1023)
    ora, gpu = str(tmp_path / "ora"), str(tmp_path / "gpu")
    os.makedirs(ora)
    os.makedirs(gpu)
    assert oracle.kmxo_build(base.encode(), 1, 1023, 7, 5, ora.encode(), None) == 0
    m = kx.get_model(1, 1023, 7, 5)
    m.init(base)
    m.save(gpu)
    import cases
    for f in ("header", "km.bin", "rest.bin"):
        assert cases.md5_file(os.path.join(gpu, f)) == cases.md5_file(os.path.join(ora, f)), f
    occ = m.kmer_to_occ(want_k)
    exact = (occ == np.minimum(want_c, 1023)).mean()
    assert exact > 0.6 and (occ == 0).mean() < 0.01          # the model answers its own k-mers (binned above 31)
    # the counter's database is also searchable the way CKMCFile::CheckKmer searches it (one bin, zero signature map),
    # and the reads it was counted from get their own k-mers' counters back (GetCountersForRead)
    db = kx.KmcDatabase(base)
    assert (db.check_kmers(want_k) == np.minimum(want_c, 1023)).all()
    with open(fq) as f:
        reads = [ln.strip() for i, ln in enumerate(f) if i % 4 == 1][:50]
    lookup = dict(zip(want_k.tolist(), np.minimum(want_c, 1023).tolist()))
    code = {"A": 0, "C": 1, "G": 2, "T": 3}
    for r, got in zip(reads, db.counters_for_reads(reads)):
        want = []
        for i in range(len(r) - 30):
            w = r[i: i + 31]
            if any(ch not in code for ch in w):
                want.append(0)
                continue
            v = rc = 0
            for ch in w:
                v = (v << 2) | code[ch]
            for ch in reversed(w):
                rc = (rc << 2) | (3 - code[ch])
            want.append(lookup.get(min(v, rc), 0))
        assert got.tolist() == want
    db.close()
