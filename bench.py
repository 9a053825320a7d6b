#!/usr/bin/env python
"""bench.py -- the kmcEx model build + kmer_to_occ path on B200, one JSON line.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]
                    [--workload hc14|rs|small|cfg1|wgs350|na12878] [--parallelism team|replicas] [--query-sweep]

A step = one model build (KModel::init: counting pass, Bloom inserts, greedy coupled-array insert, rest table) from a
synthetic KMC database of the named shape.  `value` is k-mers encoded per second with the database already resident
in HBM, `e2e` the same build through the file-based entry point (file -> pinned host -> HBM -> build; host buffers,
copies inside the timed region).  The retrieval half of the metric (kmer_to_occ queries/s) is measured in the same
run and reported under "query" (`--query-sweep`: BASELINE.json configs[4]).

Default workload: BASELINE.json configs[2] (HC14 shape, the config named for 1/2/4/8 GPUs).  N > 1: ALL ranks build
ONE model together (kmcex_b200.distributed.build_team: record range, Bloom inserts and rest sort sharded over the
ranks, coupled arrays split by ownership, exchanges through NVLink peer memory) -- strong scaling.  At N = 1 the RS
shape (configs[1], L2-resident) is measured too and reported under "extra".

Parity gate: after the timed regions the model is saved and the md5 of header / km.bin / rest.bin and of the answers
to the first 2^22 queries are compared with the UNMODIFIED reference's (tests/golden/bench_shapes.json, pinned on the
same seeded database, and/or the model oracle/_ref/ref_driver built on this box); the line carries "parity" and
the process exits non-zero on a mismatch, at every N.

`--impl reference` times the UNMODIFIED reference (oracle/_ref/ref_driver, compiled from /root/reference by
oracle/Makefile) on the host cores of this box on the same database (a capped number of builds, stated in
cpu_baseline.sample).
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
from kmcex_b200.workloads import CACHE, MODEL_FILES, SWEEP_POOL, WORKLOADS, ensure_db, golden_for, md5_file, model_digests, occ_digest  # noqa: E402,F401

PARITY_QUERIES = 1 << 22          # answers compared with the reference's (tests/golden/make_bench_golden.py: OCC_N)
REF_BUDGET_S = 150.0              # the reference arm stops starting new builds after this many seconds of CPU builds


def rank_env():
    return int(os.environ.get("RANK", 0)), int(os.environ.get("LOCAL_RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))


def config_of(workload: str, meta: dict) -> dict:
    """the `config` object: identical in both arms"""
    return {"workload": WORKLOADS[workload][4], "n_kmers": meta["n_kmers"], "k": 31, "n_hash": 7, "n_bits": 5, "ci": meta["ci"], "cs": 1023}


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


def random_sector_rates() -> dict:
    """G sectors/s the chip sustains for random 8-byte loads and random 64-bit reductions (tools/random_sector_peaks.py,
    committed under profiles/), for an L2-resident (64 MiB) and an HBM-resident (4 GiB) footprint"""
    for name in ("r2_random_sector_peaks.json", "r1_random_sector_peaks.json"):
        p = os.path.join(ROOT, "profiles", name)
        if os.path.exists(p):
            break
    else:
        return {}
    with open(p) as f:
        res = json.load(f)["results"]
    out = {"source": "profiles/" + name}
    for res_name, mb in (("l2", 64), ("hbm", 4096)):
        ld = [r["g_accesses_per_s"] for r in res if r["footprint_mb"] == mb and r["kind"] == "load8"]
        rd = [r["g_accesses_per_s"] for r in res if r["footprint_mb"] == mb and r["kind"] == "red_or64"]
        if ld and rd:
            out[res_name] = {"load": ld[0], "red": rd[0]}
    return out


def measured_traffic(workload: str, world: int):
    """dram__bytes_read.sum + dram__bytes_write.sum of insert_kernel per launch from this round's `ncu --set full` capture
    (profiles/r2_ncu_traffic.json, written by tools/ncu_summarize.py); None when there is no capture of this workload"""
    p = os.path.join(ROOT, "profiles", "r2_ncu_traffic.json")
    if world != 1 or not os.path.exists(p):
        return None
    with open(p) as f:
        return json.load(f).get(workload, {}).get("dram_bytes_per_launch")


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as f:
            return json.load(f).get("hbm_gbs", 6650.0), "measured (MEASURED_PEAKS.json hbm_gbs)"
    return 6650.0, "fallback (B200_PROFILING.md)"


# ---------------------------------------------------------------------------------------------
# the unmodified reference on the host cores: reference arm, cpu_baseline leg, parity model
# ---------------------------------------------------------------------------------------------
def ref_driver_path() -> str:
    return os.path.join(ROOT, "oracle", "_ref", "ref_driver")


def ref_model_dir(workload: str, seed: int = 1) -> str:
    return os.path.join(CACHE, f"{workload}_s{seed}", "ref_model")


def reference_build(workload: str, meta: dict) -> float:
    """one KModel::init + save of the unmodified reference; returns init seconds; stamps the model with the database digest"""
    out_dir = ref_model_dir(workload)
    os.makedirs(out_dir, exist_ok=True)
    stamp = os.path.join(out_dir, "stamp.json")
    if os.path.exists(stamp):
        os.remove(stamp)
    cores = os.cpu_count() or 1
    r = subprocess.run([ref_driver_path(), "build", meta["db"], out_dir, str(meta["ci"]), "1023", "7", "5"], capture_output=True, text=True,
                       env=dict(os.environ, OMP_NUM_THREADS=str(cores)), check=True)
    t = json.loads(r.stdout.strip().splitlines()[-1])
    with open(stamp, "w") as f:
        json.dump({"db_md5": meta["db_md5"], "model_md5": model_digests(out_dir), "init_s": t["init_s"]}, f)
    return float(t["init_s"])


def reference_stamp(workload: str, meta: dict):
    stamp = os.path.join(ref_model_dir(workload), "stamp.json")
    if not os.path.exists(stamp):
        return None
    with open(stamp) as f:
        s = json.load(f)
    return s if s.get("db_md5") == meta["db_md5"] else None


def run_reference(args) -> None:
    rank, _, world = rank_env()
    if rank != 0:
        return
    if not os.path.exists(ref_driver_path()):
        print(json.dumps({"impl": "reference", "unavailable": "oracle/_ref/ref_driver not built (needs /root/reference at build time)"}))
        return
    meta = ensure_db(args.workload)
    cores = os.cpu_count() or 1
    # a capped number of whole-database builds: warm-up builds only while they are cheap, timed builds until the budget is spent
    times, spent, warm = [], 0.0, 0
    while len(times) < max(1, args.steps):
        t = reference_build(args.workload, meta)
        spent += t
        if warm < args.warmup and t < 5.0 and spent < REF_BUDGET_S / 3:
            warm += 1
            continue
        times.append(t)
        if spent + t > REF_BUDGET_S:
            break
    ms = 1e3 * sum(times) / len(times)
    value = meta["n_kmers"] / (ms / 1e3)
    # retrieval: bounded sample of the query set, every host core (kmer_to_occ(vector, t_num), kmodel.hpp:90)
    n_q = min(meta["n_queries"], 1 << 21)
    qs = os.path.join(CACHE, f"{args.workload}_ref_q.u64")
    np.fromfile(meta["queries"], dtype=np.uint64, count=n_q).tofile(qs)
    r = subprocess.run([ref_driver_path(), "query", ref_model_dir(args.workload), qs, "31", qs + ".occ", str(cores)], capture_output=True, text=True,
                       env=dict(os.environ, OMP_NUM_THREADS=str(cores)), check=True)
    tq = json.loads(r.stdout.strip().splitlines()[-1])
    qps = n_q / tq["query_s"]
    line = {
        "impl": "reference", "metric": "kmers_encoded_per_s", "value": value, "unit": "k-mers/s", "n_gpus": args.gpus, "steps": len(times),
        "warmup": warm, "ms_per_step": ms, "higher_is_better": True, "scaling": "strong" if args.gpus > 1 else "weak", "vs_baseline": None,
        "dtype": "u64", "data": "synthetic", "config": config_of(args.workload, meta),
        "e2e": {"value": value, "unit": "k-mers/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "cpu_baseline": {"value": value, "unit": "k-mers/s", "cores": cores, "kind": "reference",
                         "sample": f"whole database ({meta['n_kmers']} k-mers), KModel::init only, {len(times)} timed build(s) after {warm} warm-up "
                                   f"(capped at {REF_BUDGET_S:.0f} s of CPU builds; requested steps={args.steps} warmup={args.warmup}); the build "
                                   "uses the reference's hard-coded 4/n_bits threads"},
        "query": {"value": qps, "unit": "queries/s", "threads": cores, "sample": f"{n_q} queries of the bench query set"},
        "gpu_launches": 0,
    }
    print(json.dumps(line))


# ---------------------------------------------------------------------------------------------
# our arm
# ---------------------------------------------------------------------------------------------
class Bench:
    def __init__(self, args):
        import torch
        import torch.distributed as dist
        import kmcex_b200 as kx
        from kmcex_b200 import distributed as kd
        self.torch, self.dist, self.kx, self.kd, self.args = torch, dist, kx, kd, args
        self.rank, self.local_rank, self.world = rank_env()
        if not torch.cuda.is_available():
            raise SystemExit("bench.py needs a CUDA device: kmcex_b200 has no CPU path")
        torch.cuda.set_device(self.local_rank)
        kx._lib.check(kx.lib().kmx_set_device(self.local_rank))
        if self.world > 1:
            os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
            dist.init_process_group("nccl", device_id=torch.device("cuda", self.local_rank))
        self.dev = torch.device("cuda", self.local_rank)
        self.flush = torch.empty(256 << 20, dtype=torch.uint8, device=self.dev)     # > 126 MB L2
        self.team = self.world > 1 and args.parallelism == "team"
        self.builds_per_step = 1 if (self.team or self.world == 1) else self.world

    def barrier(self):
        self.torch.cuda.synchronize()
        if self.world > 1:
            self.dist.barrier()
            self.torch.cuda.synchronize()

    def max_over_ranks(self, x: float) -> float:
        if self.world == 1:
            return x
        t = self.torch.tensor([x], dtype=self.torch.float64, device=self.dev)
        self.dist.all_reduce(t, op=self.dist.ReduceOp.MAX)
        return float(t.item())

    def sum_over_ranks(self, x: float) -> float:
        if self.world == 1:
            return x
        t = self.torch.tensor([x], dtype=self.torch.float64, device=self.dev)
        self.dist.all_reduce(t, op=self.dist.ReduceOp.SUM)
        return float(t.item())

    def meta_for(self, workload: str) -> dict:
        if self.rank == 0:
            ensure_db(workload)
        if self.world > 1:
            self.dist.barrier()
        return ensure_db(workload)

    def new_model(self, meta):
        return self.kx.get_model(meta["ci"], 1023, 7, 5)

    def build(self, m, db) -> None:
        """db: an opened KmcDatabase (resident or not) or a path"""
        if self.team:
            self.kd.build_team(m, db)
        else:
            m.init(db)

    # ---- build legs ------------------------------------------------------------------------
    def build_legs(self, workload: str, meta: dict, steps: int, warmup: int, sample_clocks: bool):
        kx, torch = self.kx, self.torch
        n_kmers = meta["n_kmers"]
        db = kx.KmcDatabase(meta["db"])
        if self.team:
            db.upload_share(self.rank, self.world)
        else:
            db.upload()
        infos, wall, m = [], [], None
        sampler = ClockSampler(self.local_rank) if sample_clocks else None
        launches0 = 0
        for step in range(warmup + steps):
            if step == warmup:
                if sampler:
                    sampler.start()
                launches0 = kx.lib().kmx_launch_count()
            self.flush.fill_(step & 0xFF)
            self.barrier()
            t0 = time.perf_counter()
            m = self.new_model(meta)
            self.build(m, db)
            m.sync()
            torch.cuda.synchronize()
            dt = time.perf_counter() - t0
            if step >= warmup:
                wall.append(dt)
                infos.append(m.info)
            if step < warmup + steps - 1:
                m.close()
        launches = kx.lib().kmx_launch_count() - launches0
        clocks = sampler.stop() if sampler else None
        db.close()
        t_build = self.max_over_ranks(sum(wall))
        res = {
            "n_kmers": n_kmers, "ms_per_step": 1e3 * t_build / steps, "value": self.builds_per_step * n_kmers * steps / t_build,
            "wall_ms_steps": [round(1e3 * w, 3) for w in wall], "info": infos[-1], "infos": infos, "clocks": clocks,
            "gpu_launches": int(self.sum_over_ranks(launches)),
        }
        # end to end from the files (host buffers): open, read this rank's share, copy, build
        e2e_wall = []
        for step in range(2 + steps):
            self.flush.fill_(step & 0xFF)
            self.barrier()
            t0 = time.perf_counter()
            m2 = self.new_model(meta)
            if self.team:
                db2 = kx.KmcDatabase(meta["db"])
                self.build(m2, db2)
                db2.close()
            else:
                m2.init(meta["db"])
            m2.sync()
            dt = time.perf_counter() - t0
            if step >= 2:
                e2e_wall.append(dt)
            m2.close()
        t_e2e = self.max_over_ranks(sum(e2e_wall))
        h2d = meta["suffix_bytes"] // (self.world if self.team else 1) + meta["prefix_bytes"]
        res["e2e"] = {"value": self.builds_per_step * n_kmers * steps / t_e2e, "unit": "k-mers/s",
                      "h2d_bytes_per_step": int(self.sum_over_ranks(h2d)), "d2h_bytes_per_step": 512 * self.world,
                      "ms_per_step": 1e3 * t_e2e / steps, "wall_ms_steps": [round(1e3 * w, 3) for w in e2e_wall]}
        return m, res

    # ---- retrieval -------------------------------------------------------------------------
    def query_legs(self, m, meta: dict, steps: int, warmup: int) -> dict:
        kx, torch = self.kx, self.torch
        rank, world, dev = self.rank, self.world, self.dev
        q_all = np.fromfile(meta["queries"], dtype=np.uint64)
        per = min(q_all.size // world, 1 << 24)
        q_host = torch.from_numpy(q_all[rank * per:(rank + 1) * per].astype(np.int64)).pin_memory()
        q_dev = q_host.to(dev)
        out_dev = torch.empty(per, dtype=torch.int32, device=dev)
        out_host = torch.empty(per, dtype=torch.int32).pin_memory()
        stream = torch.cuda.Stream(device=dev)                    # a real (non-NULL) stream: the kernel and the events share it
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
        q_ms = []
        for step in range(warmup + steps):
            self.flush.fill_(step & 0xFF)
            self.barrier()
            stream.wait_stream(torch.cuda.current_stream())
            ev[0].record(stream)
            m.query_device(q_dev.data_ptr(), per, out_dev.data_ptr(), stream.cuda_stream)
            ev[1].record(stream)
            torch.cuda.synchronize()
            if step >= warmup:
                q_ms.append(ev[0].elapsed_time(ev[1]))
        t_q = self.max_over_ranks(sum(q_ms) / 1e3)
        qps = world * per * steps / t_q
        q_e2e = []
        for step in range(2 + steps):
            self.flush.fill_(step & 0xFF)
            self.barrier()
            t0 = time.perf_counter()
            kx._lib.check(kx.lib().kmx_query_packed(m._h, q_host.data_ptr(), per, out_host.data_ptr()))
            dt = time.perf_counter() - t0
            if step >= 2:
                q_e2e.append(dt)
        t_qe = self.max_over_ranks(sum(q_e2e))
        qps_e2e = world * per * steps / t_qe
        assert bool((out_host.to(dev) == out_dev).all()), "host-pointer and device-pointer queries disagree"
        # the reference-facing form of the call: ASCII strings (kmer_to_occ(vector<string>)), flattened at stride k
        from kmcex_b200 import synth as _synth
        n_a = min(per, 1 << 22)
        a_host = torch.from_numpy(_synth.to_ascii(q_all[rank * per: rank * per + n_a], 31)).pin_memory()
        a_out = torch.empty(n_a, dtype=torch.int32).pin_memory()
        q_asc = []
        for step in range(2 + steps):
            self.flush.fill_(step & 0xFF)
            self.barrier()
            t0 = time.perf_counter()
            kx._lib.check(kx.lib().kmx_query_ascii(m._h, a_host.data_ptr(), 31, n_a, a_out.data_ptr()))
            dt = time.perf_counter() - t0
            if step >= 2:
                q_asc.append(dt)
        t_qa = self.max_over_ranks(sum(q_asc))
        qps_ascii = world * n_a * steps / t_qa
        assert bool((a_out == out_host[:n_a]).all()), "ASCII and packed queries disagree"
        res = {"value": qps, "unit": "queries/s", "batch": per, "ms_per_batch": 1e3 * t_q / steps,
               "e2e": {"value": qps_e2e, "unit": "queries/s", "h2d_bytes_per_step": per * 8 * world, "d2h_bytes_per_step": per * 4 * world},
               "e2e_ascii": {"value": qps_ascii, "unit": "queries/s", "batch": n_a, "h2d_bytes_per_step": n_a * 31 * world, "d2h_bytes_per_step": n_a * 4 * world,
                             "note": "kmx_query_ascii: 31-character strings at stride 31, encoded to 2 bits on the device"}}
        if self.args.query_sweep:
            res["sweep"] = self.query_sweep(m, q_all, stream, ev)
        return res

    def query_sweep(self, m, q_all, stream, ev):
        """configs[4]: 1e9 lookups in batches of 1e5 .. 1e8"""
        kx, torch = self.kx, self.torch
        rank, world, dev = self.rank, self.world, self.dev
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
            self.barrier()
            stream.wait_stream(torch.cuda.current_stream())
            for o in offs[:2]:                               # warm-up
                m.query_device(pool_dev.data_ptr() + 8 * o, B, s_out_dev.data_ptr(), stream.cuda_stream)
            ev[0].record(stream)
            for o in offs:
                m.query_device(pool_dev.data_ptr() + 8 * o, B, s_out_dev.data_ptr(), stream.cuda_stream)
            ev[1].record(stream)
            torch.cuda.synchronize()
            t_dev = self.max_over_ranks(ev[0].elapsed_time(ev[1]) / 1e3)
            n_e = min(n_b, 2000)                             # host->host: bounded number of calls for the small batches
            self.barrier()
            t0 = time.perf_counter()
            for o in offs[:n_e]:
                kx._lib.check(kx.lib().kmx_query_packed(m._h, pool_host.data_ptr() + 8 * o, B, s_out_host.data_ptr()))
            t_host = self.max_over_ranks(time.perf_counter() - t0)
            sweep.append({"batch": B, "batches": n_b, "lookups": n_b * B * world, "device_qps": world * n_b * B / t_dev,
                          "host_batches": n_e, "host_qps": world * n_e * B / t_host})
        return {"note": "BASELINE.json configs[4]: 50 % present / 37.5 % absent / 12.5 % neighbours, pool of "
                        f"{q_all.size} distinct queries walked cyclically", "unit": "queries/s", "results": sweep}

    # ---- parity gate -----------------------------------------------------------------------
    def parity(self, m, workload: str, meta: dict) -> dict:
        """md5 of the saved model and of the answers to the first PARITY_QUERIES queries, on EVERY rank, against the
        reference's digests (committed golden of the same seeded database and/or the reference model built on this box)"""
        rank = self.rank
        out_dir = os.path.join(CACHE, f"{workload}_s1", f"kmx_model_rank{rank}")
        os.makedirs(out_dir, exist_ok=True)
        # a model of several GB is saved by rank 0 only (8 x 15 GB would not fit the scratch disk): the other ranks prove that
        # their replica equals rank 0's with position-sensitive checksums of the device arrays, and answer the queries
        info = m.info
        big = self.world > 1 and (info["km_bytes"] + info["km_back_bytes"] + info["bf_bytes"]) > int(os.environ.get("KMX_BENCH_BIG_MODEL_BYTES", 4 << 30))
        sums = m.checksum()
        replicas_equal = True
        if self.world > 1:
            t = self.torch.tensor([x - (1 << 64) if x >= (1 << 63) else x for x in sums], dtype=self.torch.int64, device=self.dev)
            allt = [self.torch.empty_like(t) for _ in range(self.world)]
            self.dist.all_gather(allt, t)
            replicas_equal = all(bool((x == allt[0]).all()) for x in allt)
        saves = rank == 0 or not big
        if saves:
            m.save(out_dir)
            mine = model_digests(out_dir)
        else:
            mine = {}
        q = np.fromfile(meta["queries"], dtype=np.uint64, count=PARITY_QUERIES)
        occ_md5 = occ_digest(m.kmer_to_occ(q))
        gold = golden_for(workload)
        if gold is not None and gold["db_md5"] != meta["db_md5"]:
            gold = None                                      # another database than the pinned one (generator changed?)
        stamp = reference_stamp(workload, meta)
        res = {"checked": False, "vs": [], "db_md5_equals_golden": gold is not None}
        expect_model, expect_occ = None, None
        if gold is not None:
            expect_model, expect_occ = gold["model_md5"], (gold["occ_md5"] if gold["occ_n"] == q.size else None)
            res["vs"].append("tests/golden/bench_shapes.json (oracle/_ref on the same seeded database)")
        if stamp is not None:
            if expect_model is not None and stamp["model_md5"] != expect_model:
                res["reference_on_this_box_equals_golden"] = False
            expect_model = expect_model or stamp["model_md5"]
            res["vs"].append("oracle/_ref model built on this box")
        res["replicas_equal_on_device"] = replicas_equal
        res["files_checked_on"] = "rank 0" if big else "every rank"
        if expect_model is not None:
            res["checked"] = True
            for f in MODEL_FILES:
                res[f] = (mine[f] == expect_model[f]) if saves else replicas_equal
            if expect_occ is None and stamp is not None and rank == 0 and os.path.exists(ref_driver_path()):
                qs = os.path.join(CACHE, f"{workload}_s1", "parity_q.u64")
                q.tofile(qs)
                subprocess.run([ref_driver_path(), "query", ref_model_dir(workload), qs, "31", qs + ".occ", str(os.cpu_count() or 1)],
                               capture_output=True, text=True, check=True)
                expect_occ = occ_digest(np.fromfile(qs + ".occ", dtype=np.int32))
            if expect_occ is not None:
                res["kmer_to_occ"] = occ_md5 == expect_occ
        else:
            res["why"] = "no pinned digest for this database and no reference model on this box (run --impl reference first)"
        ok = replicas_equal and all(v for k, v in res.items() if k in MODEL_FILES or k == "kmer_to_occ")
        # KMX_BENCH_WRITE_GOLDEN=<file>: keep the reference's digests of a shape that is too large to pin in the authoring
        # container (NA12878: the reference needs ~16 min of CPU here) so that later runs can be checked without it
        dst = os.environ.get("KMX_BENCH_WRITE_GOLDEN")
        if dst and rank == 0 and stamp is not None and expect_occ is not None and ok and res["checked"]:
            entry = {"workload": WORKLOADS[workload][4], "seed": 1, "ci": meta["ci"], "n_kmers": meta["n_kmers"], "db_md5": meta["db_md5"],
                     "query_md5": meta["query_md5"], "model_md5": stamp["model_md5"], "occ_n": int(q.size), "occ_md5": expect_occ,
                     "source": "oracle/_ref/ref_driver on the GPU box's host cores (bench.py, KMX_BENCH_WRITE_GOLDEN)"}
            with open(dst, "w") as f:
                json.dump({f"{workload}_s1": entry}, f, indent=1, sort_keys=True)
        all_ok = self.sum_over_ranks(0.0 if ok else 1.0) == 0.0
        res["all_ranks"] = all_ok
        res["ranks_checked"] = self.world
        res["occ_md5"] = occ_md5
        res["model_md5"] = mine
        return res

    # ---- roofline --------------------------------------------------------------------------
    def roofline(self, workload: str, res: dict) -> dict:
        infos = res["infos"]
        info = infos[-1]
        world = self.world
        ins_ms = self.max_over_ranks(float(np.mean([i["ms_insert"] for i in infos])))
        n_active = min(world, 5) if self.team else world
        # insert_kernel (largest share of the step).  Algorithmic traffic, one 32-byte sector per touch:
        # each attempt reads n_hash cells; each accept issues n_hash cell reductions + (n_hash-2) km_back reductions.
        # Claim / reservation traffic is implementation overhead and not counted.  (team build: attempts / accepted are the
        # totals over the array owners, the time is the slowest owner's)
        loads, reds = 7 * info["insert_attempts"], 12 * info["insert_accepted"]
        if not self.team:
            loads, reds = loads * world, reds * world            # replicas: every rank does a whole build
        sectors = loads + reds
        gbs = 32.0 * sectors / (ins_ms * 1e-3) / 1e9 if ins_ms > 0 else 0.0
        peak1, peak_src = measured_peaks()
        peak = peak1 * n_active
        model_bytes = info["km_bytes"] + info["km_back_bytes"] + info["bf_bytes"]
        resident = "l2" if model_bytes < (100 << 20) else "hbm"
        rates = random_sector_rates()
        rs_peak = None
        if resident in rates and sectors:
            r = rates[resident]
            # mix-weighted: the time the measured random-access rates need for this launch's loads and reductions
            t_min = loads / (r["load"] * 1e9) + reds / (r["red"] * 1e9)
            rs_peak = 32.0 * sectors / t_min / 1e9 * n_active
        return {"kernel": "insert_kernel", "bound": "hbm", "achieved": gbs, "peak": peak, "unit": "GB/s", "frac": gbs / peak,
                "traffic": measured_traffic(workload, world), "peak_source": peak_src + (f" x {n_active} GPUs running the kernel" if n_active > 1 else ""),
                "ms_per_launch": ins_ms, "sectors_per_launch": sectors, "load_sectors": loads, "reduction_sectors": reds,
                "residency": resident, "gpus_running_the_kernel": n_active,
                "random_sector_peak_gbs": rs_peak, "frac_of_random_sector_peak": (gbs / rs_peak) if rs_peak else None,
                "random_sector_source": rates.get("source"),
                "note": "random 32-byte sectors: `peak` is the streaming-copy figure the contract asks for; the bound that applies is the "
                        "measured random-sector rate at this residency (loads and reductions weighted by this launch's mix)"
                        + ("; the model is L2-resident, so the bound is the L2's random-sector rate, not HBM" if resident == "l2" else "")}


def main() -> None:
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="kmx", choices=["kmx", "reference"])
    ap.add_argument("--workload", default="hc14", choices=sorted(WORKLOADS))
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-extra", action="store_true", help="skip the secondary RS-shape record at N = 1")
    ap.add_argument("--query-sweep", action="store_true",
                    help="BASELINE.json configs[4]: 1e9 kmer_to_occ lookups against the built model in batches of 1e5 .. 1e8 "
                         "(device-resident and host->host), reported under query.sweep")
    ap.add_argument("--parallelism", default="team", choices=["team", "replicas"],
                    help="N > 1 build: ONE model built by all ranks (kmcex_b200.distributed.build_team, strong scaling; default) or "
                         "independent whole builds per rank (weak scaling)")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
        return

    B = Bench(args)
    rank, world = B.rank, B.world
    meta = B.meta_for(args.workload)
    m, res = B.build_legs(args.workload, meta, args.steps, args.warmup, sample_clocks=True)
    query = B.query_legs(m, meta, args.steps, args.warmup)
    info = res["info"]

    line = {
        "metric": "kmers_encoded_per_s", "value": res["value"], "unit": "k-mers/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": res["ms_per_step"], "higher_is_better": True, "scaling": "strong" if (B.team or world == 1) else "weak", "vs_baseline": None,
        "dtype": "u64", "data": "synthetic", "config": config_of(args.workload, meta),
        "l2": "flushed between timed iterations (256 MiB fill)", "parallelism": (args.parallelism if world > 1 else "single"),
        "device_ms_per_step": B.max_over_ranks(float(np.mean([i["ms_total_device"] for i in res["infos"]]))),
        "wall_ms_steps": res["wall_ms_steps"],
        "stage_ms": {k: B.max_over_ranks(float(np.mean([i[k] for i in res["infos"]]))) for k in ("ms_count", "ms_encode", "ms_insert", "ms_rest")},
        "e2e": res["e2e"], "query": query, "gpu_launches": res["gpu_launches"],
        "roofline": B.roofline(args.workload, res), "clocks": res["clocks"],
        "build_stats": {k: info[k] for k in ("insert_attempts", "insert_accepted", "insert_iterations", "batches", "rest_kmers", "km_kmers", "bf_kmers", "insert_phase_cycles")},
        "model_bytes": info["km_bytes"] + info["km_back_bytes"] + info["bf_bytes"],
    }

    # ---------------- CPU baseline beside it (rank 0, N = 1): also leaves the reference model for the parity gate ----------------
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        cores = os.cpu_count() or 1
        if os.path.exists(ref_driver_path()):
            t = reference_build(args.workload, meta)
            line["cpu_baseline"] = {"value": meta["n_kmers"] / t, "unit": "k-mers/s", "cores": cores, "kind": "reference",
                                    "sample": f"whole database ({meta['n_kmers']} k-mers), KModel::init of the unmodified reference, one run"}
        else:
            line["cpu_baseline"] = {"value": None, "unit": "k-mers/s", "cores": cores, "kind": "reference", "sample": "oracle/_ref/ref_driver missing"}
    B.barrier()
    line["parity"] = B.parity(m, args.workload, meta)
    m.close()

    # ---------------- secondary record: the RS shape (configs[1], L2-resident), N = 1 only ----------------
    if world == 1 and not args.no_extra and args.workload != "rs":
        meta_rs = B.meta_for("rs")
        m_rs, r_rs = B.build_legs("rs", meta_rs, args.steps, args.warmup, sample_clocks=False)
        q_rs = B.query_legs(m_rs, meta_rs, args.steps, args.warmup)
        roof = B.roofline("rs", r_rs)
        roof["bound"] = "l2"
        line["extra"] = {"rs": {"config": config_of("rs", meta_rs), "value": r_rs["value"], "unit": "k-mers/s", "ms_per_step": r_rs["ms_per_step"],
                                "e2e": r_rs["e2e"], "query": q_rs, "roofline": roof,
                                "stage_ms": {k: float(np.mean([i[k] for i in r_rs["infos"]])) for k in ("ms_count", "ms_encode", "ms_insert", "ms_rest")},
                                "parity": B.parity(m_rs, "rs", meta_rs)}}
        m_rs.close()
    if rank == 0:
        print(json.dumps(line))
    ok = line["parity"]["all_ranks"] and (("extra" not in line) or line["extra"]["rs"]["parity"]["all_ranks"])
    if world > 1:
        B.dist.barrier()
        B.dist.destroy_process_group()
    if not ok:
        sys.stderr.write("PARITY MISMATCH: the GPU build differs from the reference's -- see \"parity\" in the line above\n")
        sys.exit(3)


if __name__ == "__main__":
    main()
