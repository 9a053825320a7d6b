"""CPU suite, part 3: the N > 1 plumbing under gloo with world_size 2 -- the team build's control plane (blob exchange,
failure protocol), contiguous sharding of the query batch, gathering in batch order, and byte-exact replication of a
model directory.  The
oracle (CPU restatement) stands in for the GPU model replica; the GPU model itself is covered by
tests/test_gpu_parity.py."""
import ctypes as C
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import cases
from kmcex_b200 import distributed as kd

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_shard_bounds_partition_the_batch():
    for n in (0, 1, 7, 8, 1000, 12345677):
        for world in (1, 2, 3, 4, 8):
            b = [kd.shard_bounds(n, r, world) for r in range(world)]
            assert b[0][0] == 0 and b[-1][1] == n
            assert all(b[i][1] == b[i + 1][0] for i in range(world - 1))
            sizes = [hi - lo for lo, hi in b]
            assert max(sizes) - min(sizes) <= 1


def _free_port() -> int:
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _load_oracle():
    lib = C.CDLL(os.path.join(ROOT, "oracle", "libkmx_oracle.so"))
    lib.kmxo_load.restype = C.c_void_p
    lib.kmxo_load.argtypes = [C.c_char_p]
    lib.kmxo_query_packed.argtypes = [C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p]
    return lib


def _team_worker(rank, world, port, base, out_dir):
    os.environ["CUDA_VISIBLE_DEVICES"] = ""               # this test is about the protocol without a device, wherever it runs
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        # the control plane of the team build: one fixed-size blob per rank, gathered in rank order
        mine = bytes([rank + 1]) * 256
        got = kd.exchange_blobs(mine)
        assert got == b"".join(bytes([r + 1]) * 256 for r in range(world))
        # the step protocol without a GPU: step 0 fails on every rank with KMX_ENOGPU, the code travels in the blobs and
        # every rank raises the same error after the same exchange -- nobody is left waiting at a barrier
        import kmcex_b200 as kx
        m = kx.get_model(1, 1023, 7, 5)
        try:
            kd.build_team(m, base)
            outcome = "built"
        except kx.KmxError as e:
            outcome = f"error {e.code}"
        with open(os.path.join(out_dir, f"outcome{rank}.txt"), "w") as f:
            f.write(outcome)
        dist.barrier()
    finally:
        dist.destroy_process_group()


def test_team_control_plane_world2(case_dbs, tmp_path):
    base, _ = case_dbs("tiny_ci1")
    world = 2
    mp.spawn(_team_worker, args=(world, _free_port(), base, str(tmp_path)), nprocs=world, join=True)
    for r in range(world):
        assert open(str(tmp_path / f"outcome{r}.txt")).read() == "error 4"     # KMX_ENOGPU, on both ranks


def test_array_ownership_covers_every_array_once():
    for n_bits in (1, 2, 5, 8):
        for n_active in range(1, n_bits + 1):
            owners = [kd.owner_of_array(a, n_active) for a in range(n_bits)]
            assert set(owners) == set(range(n_active))            # every active rank owns something
            assert max(owners.count(r) for r in range(n_active)) - min(owners.count(r) for r in range(n_active)) <= 1


def _worker(rank, world, port, model_dir, work_dir, q_path, out_path):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        # replicate the model saved by rank 0, byte for byte
        local = kd.broadcast_model_dir(model_dir if rank == 0 else None, os.path.join(work_dir, f"replica{rank}"), 0)
        lib = _load_oracle()
        h = lib.kmxo_load(local.encode())
        assert h

        def answer(q):
            q = np.ascontiguousarray(q, dtype=np.uint64)
            out = np.zeros(q.size, dtype=np.int32)
            lib.kmxo_query_packed(h, q.ctypes.data, q.size, out.ctypes.data)
            return out
        q = np.fromfile(q_path, dtype=np.uint64)
        full = kd.sharded_kmer_to_occ(answer, q)
        lo, hi, mine = kd.sharded_kmer_to_occ(answer, q, gather=False)
        assert (full[lo:hi] == mine).all()
        np.save(out_path + f".{rank}.npy", full)
    finally:
        dist.destroy_process_group()


def test_sharded_query_and_model_replication_world2(oracle, case_dbs, tmp_path):
    base, sp = case_dbs("tiny_ci1")
    model_dir = str(tmp_path / "model")
    os.makedirs(model_dir)
    assert oracle.kmxo_build(base.encode(), 1, 1023, 7, 5, model_dir.encode(), None) == 0
    q = cases.case_queries(sp)[:9001]               # odd size: uneven shards
    q_path = str(tmp_path / "q.u64")
    q.tofile(q_path)
    h = oracle.kmxo_load(model_dir.encode())
    want = np.zeros(q.size, dtype=np.int32)
    oracle.kmxo_query_packed(h, q.ctypes.data, q.size, want.ctypes.data)
    oracle.kmxo_free(h)
    out_path = str(tmp_path / "occ")
    mp.spawn(_worker, args=(2, _free_port(), model_dir, str(tmp_path), q_path, out_path), nprocs=2, join=True)
    for r in range(2):
        got = np.load(out_path + f".{r}.npy")
        assert (got == want).all()                  # identical irrespective of the rank count
        for f in kd.MODEL_FILES:
            assert cases.md5_file(os.path.join(str(tmp_path), f"replica{r}", f)) == cases.md5_file(os.path.join(model_dir, f))
