import ctypes as C
import json
import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests", "golden"))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (B200); run with -m gpu on the GPU box")


@pytest.fixture(scope="session")
def golden():
    with open(os.path.join(ROOT, "tests", "golden", "models.json")) as f:
        return json.load(f)


@pytest.fixture(scope="session")
def kat_lines():
    with open(os.path.join(ROOT, "tests", "golden", "kat.txt")) as f:
        return [ln.split() for ln in f if ln.strip()]


@pytest.fixture(scope="session")
def oracle():
    """ctypes handle of the CPU restatement (oracle/libkmx_oracle.so) -- checker only"""
    path = os.path.join(ROOT, "oracle", "libkmx_oracle.so")
    src = os.path.join(ROOT, "oracle", "kmx_oracle.cpp")
    if not os.path.exists(path) or os.path.getmtime(path) < os.path.getmtime(src):
        import subprocess
        subprocess.run(["make", "-C", os.path.join(ROOT, "oracle"), "oracle"], check=True, capture_output=True)
    lib = C.CDLL(path)
    lib.kmxo_murmur64.restype = C.c_uint64
    lib.kmxo_murmur64.argtypes = [C.c_char_p, C.c_int, C.c_uint32]
    lib.kmxo_seed.restype = C.c_uint32
    lib.kmxo_canonical.restype = C.c_uint64
    lib.kmxo_canonical.argtypes = [C.c_uint64, C.c_int]
    lib.kmxo_hash_packed.restype = C.c_uint64
    lib.kmxo_hash_packed.argtypes = [C.c_uint64, C.c_int, C.c_uint32]
    lib.kmxo_occubin.argtypes = [C.c_int, C.c_int, C.c_void_p, C.c_void_p]
    lib.kmxo_reorder.argtypes = [C.c_void_p, C.c_int, C.c_void_p]
    lib.kmxo_list.restype = C.c_int64
    lib.kmxo_list.argtypes = [C.c_char_p, C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p]
    lib.kmxo_build.argtypes = [C.c_char_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_char_p, C.c_void_p]
    lib.kmxo_load.restype = C.c_void_p
    lib.kmxo_load.argtypes = [C.c_char_p]
    lib.kmxo_free.argtypes = [C.c_void_p]
    lib.kmxo_k.argtypes = [C.c_void_p]
    lib.kmxo_query_packed.argtypes = [C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p]
    lib.kmxo_query_path.argtypes = [C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p]
    lib.kmxo_query_ascii.argtypes = [C.c_void_p, C.c_void_p, C.c_int64, C.c_int64, C.c_void_p]
    lib.kmxo_signature.restype = C.c_uint32
    lib.kmxo_signature.argtypes = [C.c_uint64, C.c_int, C.c_int]
    lib.kmxo_check_kmers.restype = C.c_int64
    lib.kmxo_check_kmers.argtypes = [C.c_char_p, C.c_void_p, C.c_int64, C.c_void_p]
    lib.kmxo_counters_for_reads.restype = C.c_int64
    lib.kmxo_counters_for_reads.argtypes = [C.c_char_p, C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p]
    return lib


@pytest.fixture(scope="session")
def ra_golden():
    with open(os.path.join(ROOT, "tests", "golden", "ra.json")) as f:
        return json.load(f)


@pytest.fixture(scope="session")
def ra_dbs(tmp_path_factory):
    """regenerates the seeded random-access databases of tests/golden/cases.py on demand (cached per session)"""
    import cases
    root = str(tmp_path_factory.mktemp("kmx_ra_cases"))
    cache = {}

    def get(name):
        if name not in cache:
            cache[name] = cases.make_ra_db(name, root)
        return cache[name]
    return get


@pytest.fixture(scope="session")
def case_dbs(tmp_path_factory):
    """regenerates the seeded databases of tests/golden/cases.py on demand (cached per session)"""
    import cases
    root = str(tmp_path_factory.mktemp("kmx_cases"))
    cache = {}

    def get(name):
        if name not in cache:
            cache[name] = cases.make_case_db(name, root)
        return cache[name]
    return get
