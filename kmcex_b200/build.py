"""Build recipe for libkmx.so (the sm_100a CUDA library behind include/kmx.h) and for the
test-only oracle.  nvcc cross-compiles without a GPU; the built .so stays in-tree so that it
travels to the GPU box with the repository snapshot."""
from __future__ import annotations

import os
import shutil
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CSRC = os.path.join(ROOT, "kmcex_b200", "csrc")
LIB = os.path.join(ROOT, "kmcex_b200", "libkmx.so")
SOURCES = ["kmx_host.cu", "kmx_team.cu", "kmx_build.cu", "kmx_sort.cu", "kmx_query.cu", "kmx_microbench.cu", "kmx_count.cu", "kmx_dist.cu", "kmx_selftest.cu", "kmx_ra.cu"]
HEADERS = ["kmx_core.cuh", "kmx_device.cuh", "kmx_launch.h", "kmx_gridbar.cuh", "kmx_internal.h", os.path.join(ROOT, "include", "kmx.h")]
NVCC_FLAGS = ["--threads", "4", "-O3", "-std=c++17", "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-Xcompiler", "-fPIC",
              "-Xcompiler", "-fvisibility=default", "-shared", "-cudart", "static", "-lz"]


def _newer(target: str, deps: list[str]) -> bool:
    if not os.path.exists(target):
        return False
    t = os.path.getmtime(target)
    return all(os.path.getmtime(d) <= t for d in deps)


def nvcc_path() -> str:
    for cand in (shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found: libkmx.so cannot be built (there is no CPU fallback)")


def build_lib(force: bool = False, verbose: bool = False, out: str | None = None, defines: list[str] | None = None) -> str:
    """out / defines: A/B builds of the kernels (e.g. defines=["KMX_INS_THREADS=512"]), loaded with KMX_LIB_PATH"""
    srcs = [os.path.join(CSRC, s) for s in SOURCES]
    deps = srcs + [h if os.path.isabs(h) else os.path.join(CSRC, h) for h in HEADERS]
    target = out or LIB
    if not force and _newer(target, deps):
        return target
    cmd = [nvcc_path()] + NVCC_FLAGS + (["-Xptxas", "-v"] if verbose else []) + [f"-D{d}" for d in (defines or [])] + ["-o", target] + srcs
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        sys.stderr.write(r.stdout + r.stderr)
        raise RuntimeError("nvcc failed building libkmx.so")
    if verbose:
        sys.stderr.write(r.stderr)
    return target


def build_oracle() -> None:
    """oracle/libkmx_oracle.so (CPU restatement) and, when /root/reference is present,
    oracle/_ref/ref_driver (the unmodified reference).  Checkers only."""
    r = subprocess.run(["make", "-C", os.path.join(ROOT, "oracle"), "all"], capture_output=True, text=True)
    if r.returncode != 0:
        sys.stderr.write(r.stdout + r.stderr)
        raise RuntimeError("building the oracle failed")


if __name__ == "__main__":
    print(build_lib(force="--force" in sys.argv, verbose="-v" in sys.argv))
    build_oracle()
