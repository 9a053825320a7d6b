#!/usr/bin/env python
"""bench.py -- the kmcEx model build + kmer_to_occ path on B200, one JSON line.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]
                    [--workload rs|small|cfg1|hc14|wgs350|na12878] [--parallelism replicas|array-owner] [--query-sweep]

A step = one model build (KModel::init: counting pass, Bloom inserts, greedy coupled-array
insert, rest table) from a synthetic KMC database of the named shape; `value` is k-mers
encoded per second with the database already resident in HBM, `e2e` the same build through
kmx_init_from_kmc (file -> pinned host -> HBM -> build, host buffers, copies inside the timed
region).  The retrieval half of the metric (kmer_to_occ queries/s) is measured in the same
run and reported under "query" (`--query-sweep`: BASELINE.json configs[4], 10^9 lookups in batches of
10^5 .. 10^8).  N > 1: see DESIGN.md "Multi-GPU" -- by default every rank builds whole models from
its own database and the query batch is sharded over the ranks (weak scaling); `--parallelism
array-owner` has all ranks build ONE model together (strong scaling).  The default workload is
BASELINE.json configs[1] (RS shape); `na12878` is configs[3], generated bin group by bin group on the GPU.

`--impl reference` times the UNMODIFIED reference (oracle/_ref/ref_driver, compiled from
/root/reference by oracle/Makefile) on the host cores of this box on the same database.
"""
from __future__ import annotations

import argparse
import ctypes as C
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

WORKLOADS = {
    # name: (synth shape, ci, lut_prefix_length, bins, BASELINE.json config it stands for)
    "small": ("small", 2, 7, 4, "test-sized (200 kbp, 40x)"),
    "cfg1": ("cfg1", 1, 3, 8, "configs[0]: 1M-read 100bp, ci1"),
    "rs": ("rs", 2, 7, 16, "configs[1]: GAGE-RS-shaped synthetic (4.6 Mbp, 100x, 101bp) k31 nh7 nb5 ci2"),
    "hc14": ("hc14", 1, 7, 64, "configs[2]: GAGE-HC14-shaped synthetic (88 Mbp, 40x) k31 nh7 nb5 ci1"),
    "wgs350": ("wgs350", 2, 7, 128, "scale check towards configs[3]: 350 Mbp synthetic genome, 30x, k31 nh7 nb5 ci2"),
    "na12878": ("na12878", 2, 7, 512, "configs[3]: NA12878-shaped synthetic (3.1 Gbp, 30x, 101bp) k31 nh7 nb5 ci2"),
}
# shapes generated bin group by bin group on the GPU (kmcex_b200.synth.make_db_streamed): (genome_bp, coverage, read_len)
STREAMED = {"na12878": (3_100_000_000, 30, 101)}
SWEEP_POOL = 100_000_000          # distinct queries behind the configs[4] sweep
CACHE = os.environ.get("KMX_BENCH_CACHE", "/tmp/kmx_bench")


def rank_env():
    return int(os.environ.get("RANK", 0)), int(os.environ.get("LOCAL_RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))


def ensure_db(workload: str, seed: int = 1):
    """generate (or reuse) the synthetic KMC database + a query set for it; returns paths and sizes"""
    from kmcex_b200 import synth
    shape, ci, lut, bins, _ = WORKLOADS[workload]
    d = os.path.join(CACHE, f"{workload}_s{seed}")
    base = os.path.join(d, "db")
    meta_path = os.path.join(d, "meta.json")
    if os.path.exists(meta_path):
        with open(meta_path) as f:
            return json.load(f)
    os.makedirs(d, exist_ok=True)
    if workload in STREAMED:
        import shutil
        import torch
        g, cov, rl = STREAMED[workload]
        need = int(g * 1.4 * 8) + (8 << 30)
        if shutil.disk_usage(d).free < need:
            raise SystemExit(f"{workload}: {need >> 30} GiB of scratch space needed under {CACHE} (set KMX_BENCH_CACHE)")
        r = synth.make_db_streamed(base, g, cov, rl, seed=seed, ci=ci, lut_prefix_length=lut, n_bins=bins, n_present=SWEEP_POOL // 2)
        torch.cuda.empty_cache()
        q = synth.mixed_queries(r["present"], SWEEP_POOL, seed=seed + 100)
        q.tofile(os.path.join(d, "queries.u64"))
        meta = {"db": base, "queries": os.path.join(d, "queries.u64"), "n_kmers": int(r["n_kmers"]), "n_queries": int(q.size), "ci": ci,
                "suffix_bytes": os.path.getsize(base + ".kmc_suf"), "prefix_bytes": os.path.getsize(base + ".kmc_pre")}
        tmp = meta_path + f".{os.getpid()}"
        with open(tmp, "w") as f:
            json.dump(meta, f)
        os.replace(tmp, meta_path)
        return meta
    sp = synth.make_db(base, shape, seed=seed, ci=ci, lut_prefix_length=lut, n_bins=bins)
    # query set: 50 % present (random strand) / 50 % absent + neighbours, BASELINE.json configs[4] mix
    n_q = 1 << 24
    q = synth.neighbour_rich_queries(sp, n_q // 2, n_q // 2 - (n_q // 2) // 4, seed=seed + 100)
    q.tofile(os.path.join(d, "queries.u64"))
    meta = {"db": base, "queries": os.path.join(d, "queries.u64"), "n_kmers": int(sp.kmers.size), "n_queries": int(q.size), "ci": ci,
            "suffix_bytes": os.path.getsize(base + ".kmc_suf"), "prefix_bytes": os.path.getsize(base + ".kmc_pre")}
    tmp = meta_path + f".{os.getpid()}"
    with open(tmp, "w") as f:
        json.dump(meta, f)
    os.replace(tmp, meta_path)
    return meta


class ClockSampler:
    """nvidia-smi clocks + throttle reasons during the timed region (B200_PROFILING.md)"""

    def __init__(self, index: int):
        self.rows, self.proc, self.index = [], None, index

    def start(self):
        q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), f"--query-gpu={q}", "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._pump, daemon=True).start()
        except OSError:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.rows.append([x.strip() for x in line.split(",")])

    def stop(self) -> dict:
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        sm, mx, reasons = [], [], set()
        for r in self.rows:
            try:
                sm.append(float(r[0]))
                mx.append(float(r[1]))
            except (ValueError, IndexError):
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None, "reasons": sorted(reasons),
                "samples": len(sm)}


# dram__bytes_read.sum + dram__bytes_write.sum of insert_kernel per launch, from profiles/ (ncu --set full)
NCU_TRAFFIC = {"rs": 9.86e8}         # profiles/r1_o_ncu_full_bench_rs.txt


def random_sector_peaks() -> dict:
    """GB/s of 32-byte sectors the chip sustains for random access (tools/random_sector_peaks.py, committed under profiles/):
    mean of the load and atomic figures, for an L2-resident (64 MiB) and an HBM-resident (4 GiB) footprint"""
    p = os.path.join(ROOT, "profiles", "r1_random_sector_peaks.json")
    if not os.path.exists(p):
        return {}
    with open(p) as f:
        res = json.load(f)["results"]
    out = {}
    for name, mb in (("l2", 64), ("hbm", 4096)):
        v = [r["gb_per_s_32B"] for r in res if r["footprint_mb"] == mb]
        if v:
            out[name] = sum(v) / len(v)
    return out


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as f:
            return json.load(f).get("hbm_gbs", 6650.0), "measured (MEASURED_PEAKS.json hbm_gbs)"
    return 6650.0, "fallback (B200_PROFILING.md)"


# ---------------------------------------------------------------------------------------------
# reference arm: the unmodified reference on the host cores
# ---------------------------------------------------------------------------------------------
def run_reference(args) -> None:
    rank, _, world = rank_env()
    if rank != 0:
        return
    ref = os.path.join(ROOT, "oracle", "_ref", "ref_driver")
    if not os.path.exists(ref):
        print(json.dumps({"impl": "reference", "unavailable": "oracle/_ref/ref_driver not built (needs /root/reference at build time)"}))
        return
    meta = ensure_db(args.workload)
    cores = os.cpu_count() or 1
    env = dict(os.environ, OMP_NUM_THREADS=str(cores))
    out_dir = os.path.join(CACHE, f"{args.workload}_ref_model")
    os.makedirs(out_dir, exist_ok=True)
    times = []
    for step in range(args.warmup + args.steps):
        r = subprocess.run([ref, "build", meta["db"], out_dir, str(meta["ci"]), "1023", "7", "5"], capture_output=True, text=True, env=env, check=True)
        t = json.loads(r.stdout.strip().splitlines()[-1])
        if step >= args.warmup:
            times.append(t["init_s"])
    ms = 1e3 * sum(times) / len(times)
    value = meta["n_kmers"] / (ms / 1e3)
    # retrieval: bounded sample of the query set, every host core (kmer_to_occ(vector, t_num), kmodel.hpp:90)
    n_q = min(meta["n_queries"], 1 << 21)
    qs = os.path.join(CACHE, f"{args.workload}_ref_q.u64")
    np.fromfile(meta["queries"], dtype=np.uint64, count=n_q).tofile(qs)
    r = subprocess.run([ref, "query", out_dir, qs, "31", qs + ".occ", str(cores)], capture_output=True, text=True, env=env, check=True)
    tq = json.loads(r.stdout.strip().splitlines()[-1])
    qps = n_q / tq["query_s"]
    line = {
        "impl": "reference", "metric": "kmers_encoded_per_s", "value": value, "unit": "k-mers/s", "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u64",
        "data": "synthetic", "config": {"workload": WORKLOADS[args.workload][4], "n_kmers": meta["n_kmers"], "k": 31, "n_hash": 7, "n_bits": 5,
                                        "ci": meta["ci"], "cs": 1023},
        "e2e": {"value": value, "unit": "k-mers/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "cpu_baseline": {"value": value, "unit": "k-mers/s", "cores": cores, "kind": "reference",
                         "sample": f"whole database ({meta['n_kmers']} k-mers), KModel::init only; build uses the reference's hard-coded 4/n_bits threads"},
        "query": {"value": qps, "unit": "queries/s", "threads": cores, "sample": f"{n_q} queries of the bench query set"},
        "gpu_launches": 0,
    }
    print(json.dumps(line))


# ---------------------------------------------------------------------------------------------
# our arm
# ---------------------------------------------------------------------------------------------
def main() -> None:
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="kmx", choices=["kmx", "reference"])
    ap.add_argument("--workload", default="rs", choices=sorted(WORKLOADS))
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--query-sweep", action="store_true",
                    help="BASELINE.json configs[4]: 1e9 kmer_to_occ lookups against the built model in batches of 1e5 .. 1e8 "
                         "(device-resident and host->host), reported under \"query_sweep\"")
    ap.add_argument("--parallelism", default="replicas", choices=["replicas", "array-owner"],
                    help="N > 1 build: independent whole builds per rank (weak scaling) or ONE build with the coupled arrays owned by "
                         "different GPUs (kmcex_b200.distributed.build_array_owner, strong scaling)")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
        return

    import torch
    import torch.distributed as dist
    import kmcex_b200 as kx

    rank, local_rank, world = rank_env()
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: kmcex_b200 has no CPU path")
    torch.cuda.set_device(local_rank)
    kx._lib.check(kx.lib().kmx_set_device(local_rank))
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    dev = torch.device("cuda", local_rank)

    if rank == 0:
        meta = ensure_db(args.workload)
    if world > 1:
        dist.barrier()
    meta = ensure_db(args.workload)
    n_kmers = meta["n_kmers"]
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)     # > 126 MB L2

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    def max_over_ranks(x: float) -> float:
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    # ---------------- build, database resident in HBM ----------------
    from kmcex_b200 import distributed as kd
    owner_mode = world > 1 and args.parallelism == "array-owner"
    builds_per_step = 1 if owner_mode else world          # array-owner: all ranks build ONE model together
    db = kx.KmcDatabase(meta["db"]).upload()
    infos, wall = [], []
    sampler = ClockSampler(local_rank)
    for step in range(args.warmup + args.steps):
        if step == args.warmup:
            sampler.start()
        flush.fill_(step & 0xFF)
        barrier()
        t0 = time.perf_counter()
        m = kx.get_model(meta["ci"], 1023, 7, 5)
        if owner_mode:
            kd.build_array_owner(m, db)
        else:
            m.init(db)
        m.sync()
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        if step >= args.warmup:
            wall.append(dt)
            infos.append(m.info)
        if step < args.warmup + args.steps - 1:
            m.close()
    clocks = sampler.stop()
    t_build = max_over_ranks(sum(wall))
    ms_per_step = 1e3 * t_build / args.steps
    value = builds_per_step * n_kmers * args.steps / t_build
    info = infos[-1]
    dev_ms = float(np.mean([i["ms_total_device"] for i in infos]))
    ins_ms = float(np.mean([i["ms_insert"] for i in infos]))
    # our kernels per build: count, tile scan, encode, one insert launch per 64 batches, rest first/index/fine/quirk (the CUB sort launches are not counted)
    launches_per_step = 3 + (info["batches"] + 63) // 64 + 4

    # ---------------- build, end to end from the files (host buffers) ----------------
    e2e_wall = []
    for step in range(2 + args.steps):
        flush.fill_(step & 0xFF)
        barrier()
        t0 = time.perf_counter()
        m2 = kx.get_model(meta["ci"], 1023, 7, 5)
        if owner_mode:
            db2 = kx.KmcDatabase(meta["db"])
            kd.build_array_owner(m2, db2)
            db2.close()
        else:
            m2.init(meta["db"])
        m2.sync()
        dt = time.perf_counter() - t0
        if step >= 2:
            e2e_wall.append(dt)
        m2.close()
    t_e2e = max_over_ranks(sum(e2e_wall))
    e2e_value = builds_per_step * n_kmers * args.steps / t_e2e

    # ---------------- retrieval ----------------
    q_all = np.fromfile(meta["queries"], dtype=np.uint64)
    per = min(q_all.size // world, 1 << 24)
    q_host = torch.from_numpy(q_all[rank * per:(rank + 1) * per].astype(np.int64)).pin_memory()
    q_dev = q_host.to(dev)
    out_dev = torch.empty(per, dtype=torch.int32, device=dev)
    out_host = torch.empty(per, dtype=torch.int32).pin_memory()
    stream = torch.cuda.Stream(device=dev)                    # a real (non-NULL) stream: the kernel and the events share it
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
    q_ms = []
    for step in range(args.warmup + args.steps):
        flush.fill_(step & 0xFF)
        barrier()
        stream.wait_stream(torch.cuda.current_stream())
        ev[0].record(stream)
        m.query_device(q_dev.data_ptr(), per, out_dev.data_ptr(), stream.cuda_stream)
        ev[1].record(stream)
        torch.cuda.synchronize()
        if step >= args.warmup:
            q_ms.append(ev[0].elapsed_time(ev[1]))
    t_q = max_over_ranks(sum(q_ms) / 1e3)
    qps = world * per * args.steps / t_q
    q_e2e = []
    for step in range(2 + args.steps):
        flush.fill_(step & 0xFF)
        barrier()
        t0 = time.perf_counter()
        kx._lib.check(kx.lib().kmx_query_packed(m._h, q_host.data_ptr(), per, out_host.data_ptr()))
        dt = time.perf_counter() - t0
        if step >= 2:
            q_e2e.append(dt)
    t_qe = max_over_ranks(sum(q_e2e))
    qps_e2e = world * per * args.steps / t_qe
    assert bool((out_host.to(dev) == out_dev).all())
    # the reference-facing form of the call: ASCII strings (kmer_to_occ(vector<string>)), flattened at stride k
    from kmcex_b200 import synth as _synth
    n_a = min(per, 1 << 22)
    a_host = torch.from_numpy(_synth.to_ascii(q_all[rank * per: rank * per + n_a], 31)).pin_memory()
    a_out = torch.empty(n_a, dtype=torch.int32).pin_memory()
    q_asc = []
    for step in range(2 + args.steps):
        flush.fill_(step & 0xFF)
        barrier()
        t0 = time.perf_counter()
        kx._lib.check(kx.lib().kmx_query_ascii(m._h, a_host.data_ptr(), 31, n_a, a_out.data_ptr()))
        dt = time.perf_counter() - t0
        if step >= 2:
            q_asc.append(dt)
    t_qa = max_over_ranks(sum(q_asc))
    qps_ascii = world * n_a * args.steps / t_qa
    assert bool((a_out == out_host[:n_a]).all())

    # ---------------- configs[4]: 1e9 lookups in batches of 1e5 .. 1e8 ----------------
    sweep = None
    if args.query_sweep:
        pool_n = q_all.size // world
        lookups = 1_000_000_000 // world                     # per rank; the batch is sharded over the ranks
        pool_host = torch.from_numpy(q_all[rank * pool_n:(rank + 1) * pool_n].astype(np.int64)).pin_memory()
        pool_dev = pool_host.to(dev)
        big = max(b for b in (100_000, 1_000_000, 10_000_000, 100_000_000) if b <= pool_n)
        s_out_dev = torch.empty(big, dtype=torch.int32, device=dev)
        s_out_host = torch.empty(big, dtype=torch.int32).pin_memory()
        sweep = []
        for B in (100_000, 1_000_000, 10_000_000, 100_000_000):
            if B > pool_n:
                continue
            n_b = max(1, lookups // B)
            span = pool_n - B + 1
            offs = [(i * B) % span for i in range(n_b)]
            barrier()
            stream.wait_stream(torch.cuda.current_stream())
            for o in offs[:2]:                               # warm-up
                m.query_device(pool_dev.data_ptr() + 8 * o, B, s_out_dev.data_ptr(), stream.cuda_stream)
            ev[0].record(stream)
            for o in offs:
                m.query_device(pool_dev.data_ptr() + 8 * o, B, s_out_dev.data_ptr(), stream.cuda_stream)
            ev[1].record(stream)
            torch.cuda.synchronize()
            t_dev = max_over_ranks(ev[0].elapsed_time(ev[1]) / 1e3)
            n_e = min(n_b, 2000)                             # host->host: bounded number of calls for the small batches
            barrier()
            t0 = time.perf_counter()
            for o in offs[:n_e]:
                kx._lib.check(kx.lib().kmx_query_packed(m._h, pool_host.data_ptr() + 8 * o, B, s_out_host.data_ptr()))
            t_host = max_over_ranks(time.perf_counter() - t0)
            sweep.append({"batch": B, "batches": n_b, "lookups": n_b * B * world, "device_qps": world * n_b * B / t_dev,
                          "host_batches": n_e, "host_qps": world * n_e * B / t_host})
        del pool_dev, s_out_dev

    # ---------------- roofline of the dominant kernel ----------------
    peak, peak_src = measured_peaks()
    rs_peaks = random_sector_peaks()
    # insert_kernel (largest share of the step).  Algorithmic traffic, one 32-byte sector per touch
    # (DESIGN.md section 3): each attempt reads n_hash cells; each accept issues n_hash cell atomics
    # + (n_hash-2) km_back atomics.  Reservation / claim traffic is implementation overhead, not counted.
    ins_sectors = 7 * info["insert_attempts"] + (7 + 5) * info["insert_accepted"]
    ins_gbs = 32.0 * ins_sectors / (ins_ms * 1e-3) / 1e9 if ins_ms > 0 else 0.0
    model_bytes = info["km_bytes"] + info["km_back_bytes"] + info["bf_bytes"]
    resident = "l2" if model_bytes < (100 << 20) else "hbm"
    rs_peak = rs_peaks.get(resident)
    roofline = {"kernel": "insert_kernel", "bound": "hbm", "achieved": ins_gbs, "peak": peak, "unit": "GB/s", "frac": ins_gbs / peak,
                "traffic": NCU_TRAFFIC.get(args.workload), "peak_source": peak_src, "ms_per_launch": ins_ms,
                "sectors_per_launch": ins_sectors, "residency": resident,
                "random_sector_peak_gbs": rs_peak, "frac_of_random_sector_peak": (ins_gbs / rs_peak) if rs_peak else None,
                "note": "random 32-byte sectors: the streaming-copy peak is not reachable by construction; the measured random-sector "
                        "peak (profiles/r1_random_sector_peaks.json: loads / atomics averaged, L2- or HBM-resident footprint) is the honest bound"}
    q_sectors_est = None
    line = {
        "metric": "kmers_encoded_per_s", "value": value, "unit": "k-mers/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "strong" if owner_mode else "weak", "vs_baseline": None, "dtype": "u64", "data": "synthetic",
        "config": {"workload": WORKLOADS[args.workload][4], "n_kmers": n_kmers, "k": 31, "n_hash": 7, "n_bits": 5, "ci": meta["ci"], "cs": 1023,
                   "l2": "flushed between timed iterations (256 MiB fill)", "parallelism": (args.parallelism if world > 1 else "single")},
        "device_ms_per_step": dev_ms, "wall_ms_steps": [round(1e3 * w, 3) for w in wall], "e2e_wall_ms_steps": [round(1e3 * w, 3) for w in e2e_wall],
        "stage_ms": {k: float(np.mean([i[k] for i in infos])) for k in ("ms_count", "ms_encode", "ms_insert", "ms_rest")},
        "e2e": {"value": e2e_value, "unit": "k-mers/s", "h2d_bytes_per_step": meta["suffix_bytes"] + meta["prefix_bytes"], "d2h_bytes_per_step": 256,
                "ms_per_step": 1e3 * t_e2e / args.steps},
        "query": {"value": qps, "unit": "queries/s", "batch": per, "ms_per_batch": 1e3 * t_q / args.steps,
                  "e2e": {"value": qps_e2e, "unit": "queries/s", "h2d_bytes_per_step": per * 8, "d2h_bytes_per_step": per * 4},
                  "e2e_ascii": {"value": qps_ascii, "unit": "queries/s", "batch": n_a, "h2d_bytes_per_step": n_a * 31, "d2h_bytes_per_step": n_a * 4,
                                "note": "kmx_query_ascii: 31-character strings at stride 31, encoded to 2 bits on the device"}},
        "gpu_launches": int(launches_per_step * args.steps),
        "roofline": roofline,
        "clocks": clocks,
        "build_stats": {k: info[k] for k in ("insert_attempts", "insert_accepted", "insert_iterations", "batches", "rest_kmers", "km_kmers", "bf_kmers", "insert_phase_cycles")},
        "model_bytes": model_bytes,
    }
    if sweep is not None:
        line["query_sweep"] = {"note": "BASELINE.json configs[4]: 50 % present / 37.5 % absent / 12.5 % neighbours, pool of "
                               f"{q_all.size} distinct queries walked cyclically; the model is far larger than the L2", "unit": "queries/s",
                               "results": sweep}

    # ---------------- CPU baseline beside it (rank 0, N = 1) ----------------
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        ref = os.path.join(ROOT, "oracle", "_ref", "ref_driver")
        cores = os.cpu_count() or 1
        if os.path.exists(ref):
            out_dir = os.path.join(CACHE, f"{args.workload}_ref_model")
            os.makedirs(out_dir, exist_ok=True)
            r = subprocess.run([ref, "build", meta["db"], out_dir, str(meta["ci"]), "1023", "7", "5"], capture_output=True, text=True,
                               env=dict(os.environ, OMP_NUM_THREADS=str(cores)))
            if r.returncode == 0:
                t = json.loads(r.stdout.strip().splitlines()[-1])
                line["cpu_baseline"] = {"value": n_kmers / t["init_s"], "unit": "k-mers/s", "cores": cores, "kind": "reference",
                                        "sample": f"whole database ({n_kmers} k-mers), KModel::init of the unmodified reference, one run"}
        if "cpu_baseline" not in line:
            line["cpu_baseline"] = {"value": None, "unit": "k-mers/s", "cores": cores, "kind": "reference", "sample": "oracle/_ref/ref_driver missing"}
    if rank == 0:
        print(json.dumps(line))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
