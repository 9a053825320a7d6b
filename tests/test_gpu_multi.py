"""GPU suite, N > 1 (skipped on a single-GPU box): ONE model built by all ranks (kmcex_b200.distributed.build_team) must be
byte-identical to the reference's on every rank for every rank count; and the replicate-then-shard query path."""
import hashlib
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import cases

pytestmark = pytest.mark.gpu


def _free_port() -> int:
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _spawn(fn, args_after_port, nprocs):
    """mp.spawn with a fresh rendezvous port; a port that got taken between the probe and the bind is retried"""
    for attempt in range(3):
        port = _free_port()
        try:
            mp.spawn(fn, args=(nprocs, port) + tuple(args_after_port), nprocs=nprocs, join=True)
            return
        except Exception as e:          # torch.multiprocessing.ProcessRaisedException carrying EADDRINUSE
            if "EADDRINUSE" not in str(e) or attempt == 2:
                raise


def _worker(rank, world, port, base, ci, work_dir, q_path):
    import kmcex_b200 as kx
    from kmcex_b200 import distributed as kd
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    kx._lib.check(kx.lib().kmx_set_device(rank))
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    try:
        def build():
            m = kx.get_model(ci, cases.MODEL["cs"], cases.MODEL["n_hash"], cases.MODEL["n_bits"])
            m.init(base)
            out = os.path.join(work_dir, "built")
            os.makedirs(out, exist_ok=True)
            m.save(out)
            return out
        sm = kd.ShardedKModel.from_builder(build, work_dir)
        q = np.fromfile(q_path, dtype=np.uint64)
        occ = sm.kmer_to_occ(q)
        np.save(os.path.join(work_dir, f"occ{rank}.npy"), occ)
    finally:
        dist.destroy_process_group()


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs 2 GPUs")
def test_two_ranks_replicated_model_sharded_queries(case_dbs, golden, tmp_path):
    name = "small_ci2"
    base, sp = case_dbs(name)
    q = cases.case_queries(sp)
    q_path = str(tmp_path / "q.u64")
    q.tofile(q_path)
    _spawn(_worker, (base, cases.CASES[name]["ci"], str(tmp_path), q_path), 2)
    for r in range(2):
        occ = np.load(str(tmp_path / f"occ{r}.npy"))
        assert hashlib.md5(occ.tobytes()).hexdigest() == golden[name]["occ_md5"]
        for f in ("header", "km.bin", "rest.bin"):
            assert cases.md5_file(str(tmp_path / f"replica_rank{r}" / f)) == golden[name]["model_md5"][f]


def _team_worker(rank, world, port, base, ci, work_dir, q_path):
    import kmcex_b200 as kx
    from kmcex_b200 import distributed as kd
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    kx._lib.check(kx.lib().kmx_set_device(rank))
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    try:
        for attempt in range(2):               # twice: the second build reuses the cached slabs and peer mappings
            m = kx.get_model(ci, cases.MODEL["cs"], cases.MODEL["n_hash"], cases.MODEL["n_bits"])
            kd.build_team(m, base)
            out = os.path.join(work_dir, f"team{attempt}_rank{rank}")
            os.makedirs(out, exist_ok=True)
            m.save(out)
            q = np.fromfile(q_path, dtype=np.uint64)
            np.save(os.path.join(work_dir, f"team{attempt}_occ{rank}.npy"), kd.ShardedKModel(m).kmer_to_occ(q))
            info = m.info
            sums = m.checksum()
            m.close()
        with open(os.path.join(work_dir, f"info{rank}.txt"), "w") as f:
            f.write(f"{info['insert_attempts']} {info['insert_accepted']} {info['rest_kmers']} {info['batches']} {sums}")
    finally:
        dist.destroy_process_group()


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs 2 GPUs")
@pytest.mark.parametrize("name,ranks", [("small_ci2", 2), ("multi_ci1", 2), ("multi_ci1", 3), ("multi_ci1", 0), ("tiny_ci1", 0)])
def test_team_build_is_byte_identical(name, ranks, case_dbs, golden, tmp_path):
    """ONE model built by all ranks: record range, Bloom inserts and rest sort sharded, coupled arrays split by ownership,
    items / survivors / finished pieces exchanged through peer memory.  Every rank (array owners and, beyond n_bits
    ranks, the ones that only decode, take a Bloom share and sort a prefix range) must end with the reference's files,
    whatever the number of ranks; ranks = 0 means every GPU of the box."""
    base, sp = case_dbs(name)
    q = cases.case_queries(sp)
    q_path = str(tmp_path / "q.u64")
    q.tofile(q_path)
    world = min(torch.cuda.device_count(), 8) if ranks == 0 else ranks
    if world > torch.cuda.device_count():
        pytest.skip(f"needs {world} GPUs")
    _spawn(_team_worker, (base, cases.CASES[name]["ci"], str(tmp_path), q_path), world)
    for attempt in range(2):
        for r in range(world):
            for f in ("header", "km.bin", "rest.bin"):
                assert cases.md5_file(str(tmp_path / f"team{attempt}_rank{r}" / f)) == golden[name]["model_md5"][f], (attempt, r, f)
            occ = np.load(str(tmp_path / f"team{attempt}_occ{r}.npy"))
            assert hashlib.md5(occ.tobytes()).hexdigest() == golden[name]["occ_md5"]
    infos = {open(str(tmp_path / f"info{r}.txt")).read() for r in range(world)}
    assert len(infos) == 1                      # every rank reports the same totals and the same device checksums
