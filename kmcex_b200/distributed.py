"""Multi-GPU plumbing (one process per GPU, torch.distributed): what SURVEY.md section 8e shards.

* retrieval: the finished model is REPLICATED on every rank (broadcast of the three model files'
  bytes from the rank that built it) and the query batch is SHARDED contiguously over the ranks;
  answers are gathered in batch order.  No per-query communication.
* build: ONE model built by all ranks (`build_team`, the default of `bench.py --gpus N`): record range, Bloom
  inserts and rest sort sharded over the ranks, coupled arrays split by ownership, every exchange through
  NVLink peer memory inside libkmx.so; torch.distributed only carries 256-byte control
  blobs between the steps.

The collectives run on CUDA tensors under NCCL and on CPU tensors under gloo (the CPU tests use
gloo with world_size 2 and the oracle standing in for the GPU model)."""
from __future__ import annotations

import os
from typing import Callable, Optional

import numpy as np
import torch
import torch.distributed as dist

MODEL_FILES = ("header", "km.bin", "rest.bin")


def shard_bounds(n: int, rank: int, world: int) -> tuple[int, int]:
    """contiguous, balanced split of range(n): rank r gets [n*r//world, n*(r+1)//world)"""
    return n * rank // world, n * (rank + 1) // world


def _comm_device(group=None) -> torch.device:
    backend = dist.get_backend(group)
    if "nccl" in str(backend):
        return torch.device("cuda", torch.cuda.current_device())
    return torch.device("cpu")


def broadcast_model_dir(src_dir: Optional[str], dst_dir: str, src_rank: int = 0, group=None) -> str:
    """Replicate a saved model (header / km.bin / rest.bin, the reference's on-disk layout) from
    src_rank to every rank: the file bytes travel as uint8 tensors (NVLink under NCCL).
    Every rank ends up with byte-identical files in dst_dir and returns that path."""
    rank = dist.get_rank(group)
    dev = _comm_device(group)
    os.makedirs(dst_dir, exist_ok=True)
    sizes = torch.zeros(len(MODEL_FILES), dtype=torch.int64, device=dev)
    if rank == src_rank:
        sizes = torch.tensor([os.path.getsize(os.path.join(src_dir, f)) for f in MODEL_FILES], dtype=torch.int64, device=dev)
    dist.broadcast(sizes, src=src_rank, group=group)
    for name, size in zip(MODEL_FILES, sizes.tolist()):
        if rank == src_rank:
            buf = torch.from_numpy(np.fromfile(os.path.join(src_dir, name), dtype=np.uint8)).to(dev)
        else:
            buf = torch.empty(size, dtype=torch.uint8, device=dev)
        if size:
            dist.broadcast(buf, src=src_rank, group=group)
        if rank != src_rank or os.path.abspath(src_dir) != os.path.abspath(dst_dir):
            buf.cpu().numpy().tofile(os.path.join(dst_dir, name))
    return dst_dir


def sharded_kmer_to_occ(answer: Callable[[np.ndarray], np.ndarray], kmers: np.ndarray, group=None, gather: bool = True):
    """kmer_to_occ over a batch every rank holds: rank r answers its contiguous shard with
    `answer` (its model replica); with gather=True every rank returns the full int32 vector in
    batch order, otherwise only (lo, hi, shard answers)."""
    rank, world = dist.get_rank(group), dist.get_world_size(group)
    n = int(kmers.shape[0])
    lo, hi = shard_bounds(n, rank, world)
    mine = np.ascontiguousarray(answer(kmers[lo:hi]), dtype=np.int32)
    if not gather:
        return lo, hi, mine
    dev = _comm_device(group)
    width = max(shard_bounds(n, r, world)[1] - shard_bounds(n, r, world)[0] for r in range(world))
    send = torch.zeros(max(width, 1), dtype=torch.int32, device=dev)
    send[: hi - lo] = torch.from_numpy(mine).to(dev)
    parts = [torch.empty_like(send) for _ in range(world)]
    dist.all_gather(parts, send, group=group)
    out = np.empty(n, dtype=np.int32)
    for r in range(world):
        a, b = shard_bounds(n, r, world)
        out[a:b] = parts[r][: b - a].cpu().numpy()
    return out


def owner_of_array(a: int, n_active: int) -> int:
    """team build: coupled array a lives on rank a % n_active (n_active = min(world, n_bits))"""
    return a % n_active


def exchange_blobs(mine: bytes, group=None) -> bytes:
    """all-gather of one fixed-size blob per rank (rank order); the control plane of the team build: counts, CUDA IPC
    handles and return codes, 256 bytes per rank and step -- bulk data never travels this way"""
    world = dist.get_world_size(group)
    dev = _comm_device(group)
    send = torch.frombuffer(bytearray(mine), dtype=torch.uint8).to(dev)
    recv = torch.empty(world * send.numel(), dtype=torch.uint8, device=dev)
    dist.all_gather_into_tensor(recv, send, group=group)
    return recv.cpu().numpy().tobytes()


def build_team(model, db, group=None) -> None:
    """KModel::init over the GPUs of one node, exactly: ONE model built by all ranks of `group` (one GPU each).

    All the work is in libkmx.so (csrc/kmx_team.cu): every rank uploads, counts and decodes only its share of the
    records; Bloom inserts are sharded and OR-ed through peer memory; array-bound k-mers go straight into the memory of
    the rank that owns their round-0 array (array a on rank a % min(world, n_bits)); the owners' insert kernels pass
    survivors on through peer memory; km_back is OR-ed, the arrays are replicated and the rest table is sorted in prefix
    ranges, one per rank -- all over NVLink peer mappings, no NCCL on the data path.  This function only carries the
    256-byte blobs (counts, IPC handles, return codes) between the steps.  Every rank ends with the complete model,
    byte-identical to a single-GPU build.  `db`: an opened KmcDatabase or the database base name."""
    from ._lib import KmxError, lib
    from .kmodel import KmcDatabase
    import ctypes as C
    rank, world = dist.get_rank(group), dist.get_world_size(group)
    own_db = not isinstance(db, KmcDatabase)
    if own_db:
        db = KmcDatabase(str(db))
    try:
        n_blob = lib().kmx_team_blob_bytes()
        blobs = None
        for step in range(lib().kmx_team_steps()):
            out = (C.c_ubyte * n_blob)()
            rc = lib().kmx_team_step(model._h, db._h, rank, world, step, blobs, out)
            msg = lib().kmx_last_error().decode(errors="replace") if rc else ""
            blobs = exchange_blobs(bytes(out), group)          # also after a failure: the codes travel in the blobs
            codes = [int.from_bytes(blobs[p * n_blob:p * n_blob + 4], "little", signed=True) for p in range(world)]
            if any(codes):
                bad = next(p for p, c in enumerate(codes) if c)
                raise KmxError(codes[bad], msg if rc else f"team build: rank {bad} failed in step {step} (code {codes[bad]})")
    finally:
        if own_db:
            db.close()


class ShardedKModel:
    """A KModel replica on this rank's GPU + the sharded batch query on top of it."""

    def __init__(self, model, group=None):
        self.model = model
        self.group = group

    @classmethod
    def from_builder(cls, build_fn: Callable[[], str], work_dir: str, src_rank: int = 0, group=None):
        """src_rank runs build_fn() -> directory of a saved model; every rank then loads a replica"""
        import kmcex_b200 as kx
        rank = dist.get_rank(group)
        src = build_fn() if rank == src_rank else None
        local = broadcast_model_dir(src, os.path.join(work_dir, f"replica_rank{rank}"), src_rank, group)
        return cls(kx.get_model(local), group)

    def kmer_to_occ(self, kmers: np.ndarray, gather: bool = True):
        return sharded_kmer_to_occ(lambda q: self.model.kmer_to_occ(np.ascontiguousarray(q, dtype=np.uint64)), kmers, self.group, gather)
