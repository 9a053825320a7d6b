"""GPU suite: the drop-in C++ boundary.  A user program written against the reference's documented
API (README.md:64-93) is compiled with g++ against include/kmodel.hpp + libkmx.so, and the `kmcEx`
command line (tools/kmcex_cli.cpp, main.cpp's flags) is run on an existing KMC database; files and
answers must equal the reference's goldens."""
import os
import shutil
import subprocess

import numpy as np
import pytest

import cases
from kmcex_b200 import synth

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIBDIR = os.path.join(ROOT, "kmcex_b200")


def _gxx():
    for c in ("/usr/bin/g++", shutil.which("g++")):
        if c and os.path.exists(c):
            return c
    pytest.skip("g++ not available")


def _compile(src, out):
    cmd = [_gxx(), "-std=c++11", "-O2", "-pthread", "-I" + os.path.join(ROOT, "include"), src, "-L" + LIBDIR, "-lkmx", "-Wl,-rpath," + LIBDIR, "-o", out]
    r = subprocess.run(cmd, capture_output=True, text=True)
    assert r.returncode == 0, r.stderr


def test_reference_style_user_program(case_dbs, golden, tmp_path):
    name = "small_ci2"
    base, sp = case_dbs(name)
    exe = str(tmp_path / "user_program")
    _compile(os.path.join(ROOT, "tests", "cpp", "user_program.cpp"), exe)
    q = cases.case_queries(sp)[:20000]
    qfile, ofile, mdir = str(tmp_path / "q.txt"), str(tmp_path / "o.txt"), str(tmp_path / "model")
    os.makedirs(mdir)
    with open(qfile, "w") as f:
        f.write("\n".join("".join(map(chr, row)) for row in synth.to_ascii(q, 31)) + "\n")
    r = subprocess.run([exe, base, mdir, qfile, ofile, str(cases.CASES[name]["ci"])], capture_output=True, text=True)
    assert r.returncode == 0, r.stdout + r.stderr
    for f in ("header", "km.bin", "rest.bin"):
        assert cases.md5_file(os.path.join(mdir, f)) == golden[name]["model_md5"][f], f
    lines = open(ofile).read().split("\n")
    got = np.array([int(x) for x in lines[: q.size]], dtype=np.int32)
    assert got[:64].tolist() == golden[name]["occ_head"]
    assert lines[q.size] == f"single {got[0]}"
    # the stdout lines of show_header_info / show_kmodel_info (kmodel.hpp:118-146)
    assert "KMCEX:" in r.stdout and "kmercount in blommfilter" in r.stdout and "kmercount hash map" in r.stdout
    assert f"total kmercount                    :     {sp.kmers.size}" in r.stdout


def test_kmcex_command_line(case_dbs, golden, tmp_path):
    name = "tiny_ci1"
    base, sp = case_dbs(name)
    exe = str(tmp_path / "kmcEx")
    _compile(os.path.join(ROOT, "tools", "kmcex_cli.cpp"), exe)
    work = str(tmp_path / "work")
    os.makedirs(work)
    r = subprocess.run([exe, "-k31", "-t4", "-ci1", "-cs1023", "-nh7", "-nb5", "reads.fastq", base, work], capture_output=True, text=True, cwd=str(tmp_path))
    assert r.returncode == 0, r.stdout + r.stderr
    save_dir = os.path.join(work, os.path.basename(base))          # main.cpp:147: workdir/basename(output)
    for f in ("header", "km.bin", "rest.bin"):
        assert cases.md5_file(os.path.join(save_dir, f)) == golden[name]["model_md5"][f], f
    # no database yet: the FASTQ input is counted on the GPU, then the model is built from it
    fq = str(tmp_path / "reads.fastq")
    want_k, want_c, _ = synth.synth_fastq(fq, genome_bp=20_000, coverage=20, read_len=100, k=31, seed=4)
    r = subprocess.run([exe, "-k31", "-ci2", fq, str(tmp_path / "counted"), work], capture_output=True, text=True, cwd=str(tmp_path))
    assert r.returncode == 0, r.stdout + r.stderr
    assert f"kept {int((want_c >= 2).sum())}" in r.stdout
    assert os.path.exists(os.path.join(work, "counted", "km.bin"))
    # too few arguments: usage + non-zero exit (main.cpp:131-134)
    r = subprocess.run([exe, "-k31"], capture_output=True, text=True)
    assert r.returncode != 0 and "kmcEx" in r.stdout
    # unopenable database: message + exit(1) (kmodel.hpp:394-397)
    r = subprocess.run([exe, "x", str(tmp_path / "nope"), work], capture_output=True, text=True, cwd=str(tmp_path))
    assert r.returncode == 1 and "can't open the kmer_data_base" in r.stdout


def test_user_program_over_two_gpus_in_one_process(case_dbs, golden, tmp_path):
    """KMX_GPUS=2: the same unmodified user program, KModel::init spread over two GPUs inside the process (one host thread
    per GPU, peer access; no NCCL, no IPC), kmer_to_occ(vector) sharded over the replicas"""
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    name = "multi_ci1"
    base, sp = case_dbs(name)
    exe = str(tmp_path / "user_program")
    _compile(os.path.join(ROOT, "tests", "cpp", "user_program.cpp"), exe)
    q = cases.case_queries(sp)
    qfile, ofile, mdir = str(tmp_path / "q.txt"), str(tmp_path / "o.txt"), str(tmp_path / "model")
    os.makedirs(mdir)
    with open(qfile, "w") as f:
        f.write("\n".join("".join(map(chr, row)) for row in synth.to_ascii(q, 31)) + "\n")
    for gpus in ("2", str(min(torch.cuda.device_count(), 8))):
        r = subprocess.run([exe, base, mdir, qfile, ofile, str(cases.CASES[name]["ci"])], capture_output=True, text=True,
                           env=dict(os.environ, KMX_GPUS=gpus))
        assert r.returncode == 0, r.stdout + r.stderr
        for f in ("header", "km.bin", "rest.bin"):
            assert cases.md5_file(os.path.join(mdir, f)) == golden[name]["model_md5"][f], (gpus, f)
        lines = open(ofile).read().split("\n")
        got = np.array([int(x) for x in lines[: q.size]], dtype=np.int32)
        import hashlib
        assert hashlib.md5(got.tobytes()).hexdigest() == golden[name]["occ_md5"]


def test_kmc_random_access_user_program(ra_dbs, ra_golden, oracle, tmp_path):
    """the CKMCFile / CKmerAPI calls of the reference's kmc_api (OpenForRA, CheckKmer, GetCountersForRead, SetMinCount ...)
    through include/kmc_ra.hpp; answers against the oracle (itself pinned to the compiled reference, tests/golden/ra.json)"""
    name = "ra_k31_sig7"
    base, sp = ra_dbs(name)
    p = cases.RA_CASES[name]
    exe = str(tmp_path / "ra_program")
    _compile(os.path.join(ROOT, "tests", "cpp", "ra_program.cpp"), exe)
    q = np.ascontiguousarray(cases.ra_queries(sp, p["seed"] + 100)[::40])
    reads = [r for r in cases.ra_reads(sp, p["seed"] + 200)[::4] + [b"ACGTN", b"A" * 31] if b"\n" not in r and r]
    kfile, rfile, ofile = str(tmp_path / "k.txt"), str(tmp_path / "r.txt"), str(tmp_path / "o.txt")
    with open(kfile, "w") as f:
        f.write("\n".join("".join(map(chr, row)) for row in synth.to_ascii(q, sp.k)) + "\n")
    with open(rfile, "wb") as f:
        f.write(b"\n".join(reads) + b"\n")
    r = subprocess.run([exe, base, kfile, rfile, ofile], capture_output=True, text=True)
    assert r.returncode == 0, r.stdout + r.stderr
    want = np.zeros(q.size, dtype=np.uint32)
    assert oracle.kmxo_check_kmers(base.encode(), q.ctypes.data, q.size, want.ctypes.data) == q.size
    flat, off = cases.flat_reads(reads)
    wr = np.zeros(int(off[-1]) + 1, dtype=np.uint32)
    n = oracle.kmxo_counters_for_reads(base.encode(), flat.ctypes.data, off.ctypes.data, len(reads), wr.ctypes.data)
    sizes = [max(0, len(x) - sp.k + 1) for x in reads]
    assert n == sum(sizes)
    cuts = np.cumsum([0] + sizes)
    lines = open(ofile).read().split("\n")
    assert lines[0] == f"k {sp.k} total {sp.kmers.size} min 1 max 1023 both 1"
    at = 1
    assert lines[at: at + q.size] == [f"{int(c != 0)} {int(c)}" for c in want]
    at += q.size
    for i, x in enumerate(reads):
        exp = "-" if len(x) < sp.k else " ".join(str(int(c)) for c in wr[cuts[i]: cuts[i + 1]])
        assert lines[at + i] == exp, (i, x)
    at += len(reads)
    assert lines[at] == "setmin 1 3 toolow 0"
    at += 1
    assert lines[at: at + q.size] == [str(int(c) if c >= 3 else 0) for c in want]          # CheckKmer under SetMinCount(3)
    at += q.size
    assert lines[at: at + q.size] == [str(int(c)) for c in want]                          # batch form after ResetMinMaxCounts
    at += q.size
    for i, x in enumerate(reads):
        exp = " ".join([str(sizes[i])] + [str(int(c)) for c in wr[cuts[i]: cuts[i + 1]]])
        assert lines[at + i] == exp, (i, x)
    assert (want != 0).sum() > 100
