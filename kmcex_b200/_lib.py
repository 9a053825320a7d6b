"""ctypes view of libkmx.so (include/kmx.h).  Loading fails loudly when the library is missing:
the package has no CPU implementation of the path."""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("KMX_LIB_PATH", os.path.join(_HERE, "libkmx.so"))   # override: A/B builds of the kernels


class KmxInfo(C.Structure):
    _fields_ = [
        ("ci", C.c_int32), ("cs", C.c_int32), ("n_hash", C.c_int32), ("n_bits", C.c_int32), ("bf_num", C.c_int32), ("k", C.c_int32),
        ("total_kmers", C.c_uint64), ("bf_kmers", C.c_uint64), ("km_kmers", C.c_uint64), ("rest_kmers", C.c_uint64),
        ("kmer_counts", C.c_uint64 * 3),
        ("bf_bytes", C.c_uint64), ("km_bytes", C.c_uint64), ("km_back_bytes", C.c_uint64), ("rest_bytes", C.c_uint64),
        ("insert_attempts", C.c_uint64), ("insert_accepted", C.c_uint64), ("insert_iterations", C.c_uint64), ("batches", C.c_uint64),
        ("insert_phase_cycles", C.c_uint64 * 12),
        ("ms_upload", C.c_float), ("ms_count", C.c_float), ("ms_encode", C.c_float), ("ms_insert", C.c_float), ("ms_rest", C.c_float),
        ("ms_total_device", C.c_float),
        ("build_time_cost", C.c_double),
    ]

    def as_dict(self) -> dict:
        out = {}
        for name, _ in self._fields_:
            v = getattr(self, name)
            out[name] = list(v) if hasattr(v, "__len__") else v
        return out


class KmxDbInfo(C.Structure):
    _fields_ = [
        ("k", C.c_uint32), ("mode", C.c_uint32), ("counter_size", C.c_uint32), ("lut_prefix_length", C.c_uint32),
        ("signature_len", C.c_uint32), ("min_count", C.c_uint32), ("max_count", C.c_uint32), ("kmc_version", C.c_uint32),
        ("total_kmers", C.c_uint64), ("lut_entries", C.c_uint64), ("suffix_bytes", C.c_uint64),
        ("record_bytes", C.c_uint32), ("on_device", C.c_int32), ("both_strands", C.c_uint32),
    ]

    def as_dict(self) -> dict:
        return {name: getattr(self, name) for name, _ in self._fields_}


class KmxCountInfo(C.Structure):
    _fields_ = [("n_reads", C.c_uint64), ("n_windows", C.c_uint64), ("n_unique", C.c_uint64), ("n_kept", C.c_uint64),
                ("lut_prefix_length", C.c_uint32), ("counter_size", C.c_uint32)]


# name -> (restype, argtypes); kept in one table so that the CPU test can check it against kmx.h
SIGNATURES = {
    "kmx_last_error": (C.c_char_p, []),
    "kmx_device_count": (C.c_int, []),
    "kmx_set_device": (C.c_int, [C.c_int]),
    "kmx_version": (C.c_char_p, []),
    "kmx_create": (C.c_void_p, [C.c_int, C.c_int, C.c_int, C.c_int]),
    "kmx_load": (C.c_void_p, [C.c_char_p]),
    "kmx_destroy": (None, [C.c_void_p]),
    "kmx_init_from_kmc": (C.c_int, [C.c_void_p, C.c_char_p]),
    "kmx_init_from_db": (C.c_int, [C.c_void_p, C.c_void_p]),
    "kmx_save": (C.c_int, [C.c_void_p, C.c_char_p]),
    "kmx_query_ascii": (C.c_int, [C.c_void_p, C.c_void_p, C.c_size_t, C.c_size_t, C.c_void_p]),
    "kmx_query_packed": (C.c_int, [C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p]),
    "kmx_query_packed_device": (C.c_int, [C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p, C.c_void_p]),
    "kmx_query_ascii_device": (C.c_int, [C.c_void_p, C.c_void_p, C.c_size_t, C.c_size_t, C.c_void_p, C.c_void_p]),
    "kmx_query_path_packed": (C.c_int, [C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p]),
    "kmx_info": (None, [C.c_void_p, C.POINTER(KmxInfo)]),
    "kmx_model_sync": (C.c_int, [C.c_void_p]),
    "kmx_db_open": (C.c_void_p, [C.c_char_p]),
    "kmx_db_upload": (C.c_int, [C.c_void_p]),
    "kmx_db_info": (None, [C.c_void_p, C.POINTER(KmxDbInfo)]),
    "kmx_db_list": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.POINTER(C.c_uint64)]),
    "kmx_db_close": (None, [C.c_void_p]),
    "kmx_db_check_kmers": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p]),
    "kmx_db_set_count_range": (C.c_int, [C.c_void_p, C.c_uint32, C.c_uint32]),
    "kmx_db_reset_count_range": (C.c_int, [C.c_void_p]),
    "kmx_db_counters_for_reads": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p, C.POINTER(C.c_int64)]),
    "kmx_host_murmur64": (C.c_uint64, [C.c_char_p, C.c_int, C.c_uint32]),
    "kmx_host_hash_packed": (C.c_uint64, [C.c_uint64, C.c_int, C.c_uint32]),
    "kmx_host_canonical": (C.c_uint64, [C.c_uint64, C.c_int]),
    "kmx_host_seed": (C.c_uint32, [C.c_int]),
    "kmx_host_occubin": (C.c_int, [C.c_int, C.c_int, C.c_void_p, C.c_void_p]),
    "kmx_host_sizes": (None, [C.POINTER(C.c_uint64), C.c_int, C.c_uint64, C.c_int, C.POINTER(C.c_uint64)]),
    "kmx_host_fastmod": (C.c_uint64, [C.c_uint64, C.c_uint64]),
    "kmx_host_signature": (C.c_uint32, [C.c_uint64, C.c_int, C.c_int]),
    "kmx_host_reorder": (C.c_int, [C.c_void_p, C.c_int, C.c_void_p]),
    "kmx_count_fastq": (C.c_int, [C.POINTER(C.c_char_p), C.c_int, C.c_int, C.c_int, C.c_int, C.c_char_p, C.POINTER(KmxCountInfo)]),
    "kmx_set_devices": (C.c_int, [C.POINTER(C.c_int), C.c_int]),
    "kmx_db_upload_share": (C.c_int, [C.c_void_p, C.c_int, C.c_int]),
    "kmx_team_steps": (C.c_int, []),
    "kmx_team_blob_bytes": (C.c_int, []),
    "kmx_team_step": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p]),
    "kmx_launch_count": (C.c_ulonglong, []),
    "kmx_model_checksum": (C.c_int, [C.c_void_p, C.POINTER(C.c_uint64)]),
    "kmx_selftest_positions": (C.c_int, [C.c_void_p, C.c_size_t, C.c_int, C.c_uint64, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p]),
    "kmx_host_route": (None, [C.c_uint64, C.c_int, C.c_int, C.POINTER(C.c_int32), C.POINTER(C.c_uint64)]),
    "kmx_host_prefix_cuts": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_void_p]),
    "kmx_microbench_random": (C.c_int, [C.c_int, C.c_uint64, C.c_uint64, C.c_int, C.POINTER(C.c_float)]),
    "kmx_microbench_grid_barrier": (C.c_int, [C.c_int, C.c_int, C.c_int, C.c_int, C.POINTER(C.c_float)]),
    "kmx_microbench_hot_atomic": (C.c_int, [C.c_int, C.c_int, C.POINTER(C.c_float)]),
    "kmx_microbench_peer_random": (C.c_int, [C.c_int, C.c_int, C.c_int, C.c_uint64, C.c_uint64, C.c_int, C.POINTER(C.c_float)]),
    "kmx_microbench_stream_read": (C.c_int, [C.c_uint64, C.c_int, C.c_int, C.POINTER(C.c_float)]),
    "kmx_microbench_windowed": (C.c_int, [C.c_int, C.c_uint64, C.c_uint64, C.c_uint64, C.c_int, C.POINTER(C.c_float)]),
}

_lib = None


def lib() -> C.CDLL:
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(
                f"{LIB_PATH} is missing: build it with `python -m kmcex_b200.build` (nvcc, sm_100a). "
                "kmcex_b200 has no CPU implementation of the build/query path.")
        _lib = C.CDLL(LIB_PATH)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(_lib, name)
            fn.restype = res
            fn.argtypes = args
    return _lib


class KmxError(RuntimeError):
    def __init__(self, code: int, message: str):
        super().__init__(f"kmx error {code}: {message}")
        self.code = code


def check(rc: int) -> None:
    if rc != 0:
        raise KmxError(rc, lib().kmx_last_error().decode(errors="replace"))
