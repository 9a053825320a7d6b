"""Python mirror of the reference's KModel API (kmodel.hpp) on top of the C ABI of libkmx.so.

    get_model(ci, cs, num_hash, num_bit) / get_model(save_dir)      kmodel.hpp:674,680
    KModel.init / init_KModel(db_file)                              kmodel.hpp:57, README.md:76
    KModel.save / save_model(dir)                                   kmodel.hpp:173, README.md:78
    KModel.kmer_to_occ(str | list[str] | ndarray)                   kmodel.hpp:90-116

Everything computes on the GPU through libkmx.so; there is no CPU path in this package."""
from __future__ import annotations

import ctypes as C

import numpy as np

from ._lib import KmxDbInfo, KmxError, KmxInfo, check, lib


class KmcDatabase:
    """A KMC database opened for listing (CKMCFile::OpenForListing, kmc_file.cpp:66-99)."""

    def __init__(self, db_base: str):
        self._h = lib().kmx_db_open(db_base.encode())
        if not self._h:
            raise KmxError(2, lib().kmx_last_error().decode(errors="replace"))

    def upload(self) -> "KmcDatabase":
        check(lib().kmx_db_upload(self._h))
        return self

    def upload_share(self, rank: int, world: int) -> "KmcDatabase":
        """only the tile range a team build's rank decodes (kmcex_b200.distributed.build_team)"""
        check(lib().kmx_db_upload_share(self._h, rank, world))
        return self

    @property
    def info(self) -> dict:
        i = KmxDbInfo()
        lib().kmx_db_info(self._h, C.byref(i))
        return i.as_dict()

    def list(self) -> tuple[np.ndarray, np.ndarray]:
        """(packed k-mers, counts) in listing order (ReadNextKmer, kmc_file.cpp:428-515), decoded on the GPU"""
        total = self.info["total_kmers"]
        kmers = np.empty(max(total, 1), dtype=np.uint64)
        counts = np.empty(max(total, 1), dtype=np.uint32)
        n = C.c_uint64(0)
        check(lib().kmx_db_list(self._h, kmers.ctypes.data, counts.ctypes.data, C.byref(n)))
        return kmers[: n.value], counts[: n.value]

    # ---- random access (CKMCFile::OpenForRA / CheckKmer / GetCountersForRead, kmc_file.cpp:27-58,320-356,879-897) ----
    def check_kmers(self, kmers) -> np.ndarray:
        """exact counter of every packed k-mer (as given, not canonicalised): 0 when absent or outside [min_count, max_count]"""
        q = np.ascontiguousarray(kmers, dtype=np.uint64)
        out = np.zeros(q.size, dtype=np.uint32)
        check(lib().kmx_db_check_kmers(self._h, q.ctypes.data, q.size, out.ctypes.data))
        return out

    def counters_for_reads(self, reads) -> list[np.ndarray]:
        """one array of len(read) - k + 1 counters per read (empty for reads shorter than k); windows holding a byte other
        than ACGTacgt count 0; canonical k-mers are looked up when the database holds both strands"""
        raw = [r.encode() if isinstance(r, str) else bytes(r) for r in reads]
        offsets = np.zeros(len(raw) + 1, dtype=np.int64)
        np.cumsum([len(r) for r in raw], out=offsets[1:])
        flat = np.frombuffer(b"".join(raw) or b"\0", dtype=np.uint8)
        k = self.info["k"]
        sizes = [max(0, len(r) - k + 1) for r in raw]
        out = np.zeros(max(1, sum(sizes)), dtype=np.uint32)
        n = C.c_int64(0)
        check(lib().kmx_db_counters_for_reads(self._h, flat.ctypes.data, offsets.ctypes.data, len(raw), out.ctypes.data, C.byref(n)))
        assert n.value == sum(sizes)
        cuts = np.cumsum([0] + sizes)
        return [out[cuts[i]: cuts[i + 1]] for i in range(len(raw))]

    def close(self) -> None:
        if self._h:
            lib().kmx_db_close(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class KModel:
    def __init__(self, handle):
        self._h = handle

    # ---- build -------------------------------------------------------------------------
    def init(self, db_file) -> None:
        if isinstance(db_file, KmcDatabase):
            check(lib().kmx_init_from_db(self._h, db_file._h))
        else:
            check(lib().kmx_init_from_kmc(self._h, str(db_file).encode()))

    init_KModel = init

    def save(self, save_dir: str) -> None:
        check(lib().kmx_save(self._h, str(save_dir).encode()))

    save_model = save

    # ---- retrieval ---------------------------------------------------------------------
    def kmer_to_occ(self, kmers, t_num: int = 4):
        """str -> int; sequence of str / (n,k) uint8 ASCII matrix / uint64 packed array -> int32 array.
        t_num is accepted for signature compatibility (kmodel.hpp:90) and ignored: the batch runs on the GPU."""
        if isinstance(kmers, (str, bytes)):
            return int(self.kmer_to_occ([kmers])[0])
        if isinstance(kmers, np.ndarray) and kmers.dtype == np.uint64:
            q = np.ascontiguousarray(kmers)
            out = np.empty(q.size, dtype=np.int32)
            check(lib().kmx_query_packed(self._h, q.ctypes.data, q.size, out.ctypes.data))
            return out
        if isinstance(kmers, np.ndarray) and kmers.dtype == np.uint8 and kmers.ndim == 2:
            flat = np.ascontiguousarray(kmers)
            n, stride = flat.shape
        else:
            k = self.info["k"]
            if k == 0:      # not initialised: let the library report it
                check(lib().kmx_query_packed(self._h, None, 0, None))
            rows =[s.encode() if isinstance(s, str) else bytes(s) for s in kmers]
            if any(len(r) != k for r in rows):
                raise ValueError(f"every k-mer must have length k={k} (the reference answers 0 or garbage otherwise, rest.hpp:224)")
            flat = np.frombuffer(b"".join(rows), dtype=np.uint8).reshape(len(rows), k) if rows else np.zeros((0, k), np.uint8)
            n, stride = flat.shape
        out = np.empty(n, dtype=np.int32)
        check(lib().kmx_query_ascii(self._h, flat.ctypes.data, stride, n, out.ctypes.data))
        return out

    def query_device(self, d_kmers_ptr: int, n: int, d_out_ptr: int, stream: int = 0) -> None:
        """device-resident packed query (pointers on the model's device), asynchronous"""
        check(lib().kmx_query_packed_device(self._h, d_kmers_ptr, n, d_out_ptr, stream))

    def query_path(self, kmers: np.ndarray) -> np.ndarray:
        q = np.ascontiguousarray(kmers, dtype=np.uint64)
        out = np.empty(q.size, dtype=np.int32)
        check(lib().kmx_query_path_packed(self._h, q.ctypes.data, q.size, out.ctypes.data))
        return out

    def checksum(self) -> list:
        """four position-sensitive 64-bit checksums of the device arrays (filters, coupled arrays, rest keys, rest counts + index)"""
        sums = (C.c_uint64 * 4)()
        check(lib().kmx_model_checksum(self._h, sums))
        return [int(x) for x in sums]

    def sync(self) -> None:
        check(lib().kmx_model_sync(self._h))

    # ---- reporting (kmodel.hpp:118-169) ---------------------------------------------------
    @property
    def info(self) -> dict:
        i = KmxInfo()
        lib().kmx_info(self._h, C.byref(i))
        return i.as_dict()

    def show_header_info(self) -> None:
        i = self.info
        print("KMCEX:")
        print(f"   kmodel number hash                 :     {i['n_hash']}")
        print(f"   kmodel bit array                   :     {i['n_bits']}")
        print(f"   total kmercount                    :     {i['total_kmers']}")
        print(f"   kmercount in blommfilter           :     {i['bf_kmers']}")
        print(f"   kmercount in kmodel                :     {i['km_kmers']}")

    def show_kmodel_info(self) -> None:
        i = self.info
        mb = lambda b: f"{b // (1024 * 1024)}MB"
        total = i["bf_bytes"] + i["km_bytes"] + i["rest_bytes"] + i["km_back_bytes"]
        print(f"   kmercount hash map                 :     {i['rest_kmers']}")
        print(f"   memory bloomfilter                 :     {mb(i['bf_bytes'])}")
        print(f"   memory bit array                   :     {mb(i['km_bytes'])}")
        print(f"   memory rest map                    :     {mb(i['rest_bytes'])}")
        print(f"   total memory                       :     {mb(total)}")
        print(f"   build time cost                    :     {i['build_time_cost']:g}")

    def close(self) -> None:
        if self._h:
            lib().kmx_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def get_model(ci_or_dir=1, cs: int = 1023, num_hash: int = 7, num_bit: int = 5) -> KModel:
    """get_model(ci, cs, num_hash, num_bit) or get_model(save_dir) -- kmodel.hpp:674-696"""
    if isinstance(ci_or_dir, (str, bytes)):
        d = ci_or_dir.encode() if isinstance(ci_or_dir, str) else ci_or_dir
        h = lib().kmx_load(d)
    else:
        h = lib().kmx_create(int(ci_or_dir), int(cs), int(num_hash), int(num_bit))
    if not h:
        raise KmxError(1, lib().kmx_last_error().decode(errors="replace"))
    return KModel(h)
