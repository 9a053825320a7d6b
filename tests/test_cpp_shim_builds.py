"""CPU suite: the drop-in C++ boundary compiles and links without a GPU.  A translation unit written against the
reference's documented API (README.md:64-93) and the `kmcEx` command line (main.cpp's flags) build with plain g++
against include/kmodel.hpp + libkmx.so; without a device the model build must fail loudly (message + exit code 1,
the reference's own error convention, kmodel.hpp:394-397), never fall back to a host path."""
import os
import shutil
import subprocess

import pytest

import kmcex_b200 as kx

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIBDIR = os.path.join(ROOT, "kmcex_b200")


def _compile(src, out):
    gxx = shutil.which("g++")
    if not gxx:
        pytest.skip("g++ not available")
    cmd = [gxx, "-std=c++11", "-O2", "-pthread", "-Wall", "-I" + os.path.join(ROOT, "include"), src, "-L" + LIBDIR, "-lkmx",
           "-Wl,-rpath," + LIBDIR, "-o", out]
    r = subprocess.run(cmd, capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    return r.stderr


def test_user_program_and_cli_build_against_the_shim(tmp_path):
    _compile(os.path.join(ROOT, "tests", "cpp", "user_program.cpp"), str(tmp_path / "user_program"))
    _compile(os.path.join(ROOT, "tools", "kmcex_cli.cpp"), str(tmp_path / "kmcex_cli"))
    r = subprocess.run([str(tmp_path / "kmcex_cli")], capture_output=True, text=True)
    assert r.returncode != 0 and "kmcEx" in (r.stdout + r.stderr)          # usage text, as main.cpp prints it


def test_cli_fails_loudly_without_a_gpu(case_dbs, tmp_path):
    if kx.lib().kmx_device_count() > 0:
        pytest.skip("a GPU is present: the GPU suite runs the command line for real")
    _compile(os.path.join(ROOT, "tools", "kmcex_cli.cpp"), str(tmp_path / "kmcex_cli"))
    base, _ = case_dbs("small_ci2")
    work = tmp_path / "work"
    work.mkdir()
    r = subprocess.run([str(tmp_path / "kmcex_cli"), "-k31", "-ci2", "reads.fq", base, str(work)], capture_output=True, text=True, cwd=str(tmp_path))
    assert r.returncode == 1, (r.returncode, r.stdout, r.stderr)
    assert "CUDA" in r.stdout + r.stderr or "GPU" in r.stdout + r.stderr or "device" in r.stdout + r.stderr
    assert not (work / os.path.basename(base) / "km.bin").exists()


def test_random_access_program_builds_and_fails_loudly_without_a_gpu(ra_dbs, tmp_path):
    """include/kmc_ra.hpp (CKMCFile / CKmerAPI names of the reference's kmc_api): compiles with -Wall, and OpenForRA
    reports failure -- it does not fall back to a host search -- when there is no device"""
    warnings = _compile(os.path.join(ROOT, "tests", "cpp", "ra_program.cpp"), str(tmp_path / "ra_program"))
    assert "warning" not in warnings
    if kx.lib().kmx_device_count() > 0:
        pytest.skip("a GPU is present: the GPU suite runs the program for real")
    base, _ = ra_dbs("ra_k23_one_strand")
    for name in ("k.txt", "r.txt"):
        (tmp_path / name).write_text("ACGT\n")
    r = subprocess.run([str(tmp_path / "ra_program"), base, str(tmp_path / "k.txt"), str(tmp_path / "r.txt"), str(tmp_path / "o.txt")], capture_output=True, text=True)
    assert r.returncode == 1 and "cannot open" in r.stdout
    assert "CUDA" in r.stdout or "GPU" in r.stdout or "device" in r.stdout
