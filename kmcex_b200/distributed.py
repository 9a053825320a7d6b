"""Multi-GPU plumbing (one process per GPU, torch.distributed): what SURVEY.md section 8e shards.

* retrieval: the finished model is REPLICATED on every rank (broadcast of the three model files'
  bytes from the rank that built it) and the query batch is SHARDED contiguously over the ranks;
  answers are gathered in batch order.  No per-query communication.
* build: either every rank builds whole models (independent databases -- `bench.py --gpus N`), or
  ONE model is built by all ranks (`build_array_owner`): Bloom inserts sharded by record range and
  OR-ed through peer memory, coupled arrays split by ownership with survivors handed from owner to
  owner through peer memory.

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


class _DeviceBytes:
    """a raw device allocation of libkmx.so as something torch can wrap without copying"""

    def __init__(self, ptr: int, nbytes: int):
        self.__cuda_array_interface__ = {"shape": (int(nbytes),), "typestr": "|u1", "data": (int(ptr), False), "version": 2}


def _view(ptr: int, nbytes: int, dtype=torch.uint8) -> torch.Tensor:
    t = torch.as_tensor(_DeviceBytes(ptr, nbytes), device=torch.device("cuda", torch.cuda.current_device()))
    return t.view(dtype)


def or_merge_(buf: torch.Tensor, group=None) -> torch.Tensor:
    """bitwise OR of `buf` over all ranks, in place (NCCL has no OR reduction: all-gather + local OR)"""
    world = dist.get_world_size(group)
    if world == 1:
        return buf
    parts = [torch.empty_like(buf) for _ in range(world)]
    dist.all_gather(parts, buf, group=group)
    acc = parts[0]
    for p in parts[1:]:
        acc |= p
    buf.copy_(acc)
    return buf


def concat_ranks(local: torch.Tensor, n_local: int, group=None) -> torch.Tensor:
    """concatenation over ranks (rank order) of the first n_local elements of `local`; every rank gets the whole"""
    world = dist.get_world_size(group)
    dev = local.device
    counts = torch.zeros(world, dtype=torch.int64, device=dev)
    counts[dist.get_rank(group)] = n_local
    dist.all_reduce(counts, group=group)
    counts = counts.tolist()
    width = max(max(counts), 1)
    send = torch.zeros(width, dtype=local.dtype, device=dev)
    send[:n_local] = local[:n_local]
    parts = [torch.empty_like(send) for _ in range(world)]
    dist.all_gather(parts, send, group=group)
    return torch.cat([p[:c] for p, c in zip(parts, counts)]) if sum(counts) else torch.zeros(0, dtype=local.dtype, device=dev)


def owner_of_array(a: int, n_active: int) -> int:
    """array-owner decomposition: coupled array a lives on rank a % n_active"""
    return a % n_active


def build_array_owner(model, db, group=None, n_active: Optional[int] = None) -> None:
    """KModel::init over the GPUs of one node, exactly (option A).

    Every rank runs the counting pass and inserts the Bloom-bound records of its share of the
    database; the partial Bloom filters are OR-ed over all ranks through peer memory (libkmx's own
    kernel, kmx_dist_merge).  The coupled arrays are owned round-robin by ranks 0..n_active-1
    (n_active <= n_bits), whose insert kernels pass survivors to the next owner through peer-mapped
    memory; km_back is OR-ed the same way as the filters.  Afterwards the owned arrays are broadcast
    and the survivor lists are concatenated, so that every rank holds the complete model --
    byte-identical to a single-GPU build."""
    from ._lib import KmxDistBuffers, check, lib
    import ctypes as C
    rank, world = dist.get_rank(group), dist.get_world_size(group)
    n_bits = model.info["n_bits"]
    n_active = min(world, n_bits, 8) if n_active is None else n_active
    dev = torch.device("cuda", torch.cuda.current_device())
    handles = (C.c_ubyte * 128)()
    check(lib().kmx_dist_prepare(model._h, db._h, rank, n_active, world, handles))
    mine = torch.tensor(list(handles), dtype=torch.uint8, device=dev)
    allh = [torch.empty_like(mine) for _ in range(world)]
    dist.all_gather(allh, mine, group=group)
    blob = bytes(torch.cat(allh).cpu().numpy().tobytes())
    check(lib().kmx_dist_connect(model._h, blob))
    dist.barrier(group=group)                      # every rank has mapped its peers before any kernel writes
    check(lib().kmx_dist_merge(model._h, 0))       # Bloom filters: OR of the ranks' shares
    check(lib().kmx_dist_insert(model._h))
    check(lib().kmx_dist_merge(model._h, 1))       # km_back: OR of what every array owner accepted
    bufs = KmxDistBuffers()
    check(lib().kmx_dist_buffers(model._h, C.byref(bufs)))
    check(lib().kmx_model_sync(model._h))          # the collectives below run on torch's stream
    for a in range(n_bits):                        # owners publish their arrays
        dist.broadcast(_view(bufs.cells[a], bufs.cell_bytes), src=owner_of_array(a, n_active), group=group)
    n_local = int(bufs.rest_n)
    cap = max(n_local, 1)
    rest_k = concat_ranks(_view(bufs.rest_kmer, cap * 8, torch.int64) if bufs.rest_kmer else torch.zeros(1, dtype=torch.int64, device=dev), n_local, group)
    rest_o = concat_ranks(_view(bufs.rest_occ, cap * 4, torch.int32) if bufs.rest_occ else torch.zeros(1, dtype=torch.int32, device=dev), n_local, group)
    stats = torch.tensor([bufs.insert_attempts, bufs.insert_accepted], dtype=torch.int64, device=dev)
    dist.all_reduce(stats, group=group)
    torch.cuda.synchronize()
    check(lib().kmx_dist_finish(model._h, rest_k.data_ptr() if rest_k.numel() else None, rest_o.data_ptr() if rest_o.numel() else None,
                                rest_k.numel(), int(stats[0]), int(stats[1])))


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
