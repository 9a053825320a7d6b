"""FASTQ -> KMC database on the GPU: the stage the reference delegates to the external `kmc` binary
(main.cpp:136-140: kmc -k -t -ci -cs input output tmp).  Thin wrapper of kmx_count_fastq."""
from __future__ import annotations

import ctypes as C

from ._lib import KmxCountInfo, check, lib


def count_fastq(paths, out_base: str, k: int = 31, ci: int = 1, cs: int = 1023) -> dict:
    """paths: one FASTQ file (plain text or gzip) or a list of them.  Writes <out_base>.kmc_pre/.kmc_suf."""
    if isinstance(paths, (str, bytes)):
        paths = [paths]
    arr = (C.c_char_p * len(paths))(*[p.encode() if isinstance(p, str) else p for p in paths])
    info = KmxCountInfo()
    check(lib().kmx_count_fastq(arr, len(paths), k, ci, cs, out_base.encode(), C.byref(info)))
    return {name: getattr(info, name) for name, _ in info._fields_}
