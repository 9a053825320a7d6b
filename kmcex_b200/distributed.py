"""Multi-GPU plumbing (one process per GPU, torch.distributed): what SURVEY.md section 8e shards.

* retrieval: the finished model is REPLICATED on every rank (broadcast of the three model files'
  bytes from the rank that built it) and the query batch is SHARDED contiguously over the ranks;
  answers are gathered in batch order.  No per-query communication.
* build: in this round every rank builds whole models (independent databases, or replicas of the
  same one) -- the coupled-array insert is sequential per array (kmodel.hpp:557-573) and its
  array-owner decomposition is the next row.

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
