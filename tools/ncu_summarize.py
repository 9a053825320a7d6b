#!/usr/bin/env python
"""Turn ncu outputs brought back in gpurun_out/ into the text summaries kept under profiles/.
  ncu_summarize.py launches <launches.csv> <out.txt>     per-kernel totals and shares of a launch list
  ncu_summarize.py full <report.ncu-rep> <out.txt>       key metrics of every captured launch of a --set full report
  ncu_summarize.py traffic <report.ncu-rep> <workload> <kernel substring> <profiles/r2_ncu_traffic.json>
                                                         mean dram read + write bytes per launch of that kernel -> the json bench.py reads"""
import collections
import csv
import subprocess
import sys


def launches(src, dst):
    lines = [l for l in open(src) if not l.startswith("==")]
    agg = collections.OrderedDict()
    for row in csv.DictReader(lines):
        v = float(row["Metric Value"].replace(",", ""))
        u = row["Metric Unit"]
        v = v / 1e3 if u == "ns" else v * 1e3 if u == "ms" else v
        agg.setdefault(row["Kernel Name"][:70], []).append(v)
    tot = sum(sum(v) for v in agg.values())
    with open(dst, "w") as f:
        f.write(f"# source: {src}; gpu__time_duration.sum per launch, --clock-control none (cold-cache, serialised: compare shares)\n")
        for k, v in agg.items():
            f.write(f"{k:72s} n={len(v):4d} sum_us={sum(v):11.1f} mean_us={sum(v) / len(v):10.1f} share={100 * sum(v) / tot:5.1f}%\n")
    print(open(dst).read())


KEYS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "lts__t_sectors.sum", "lts__t_sector_hit_rate.pct",
        "l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum", "l1tex__t_sectors_pipe_lsu_mem_global_op_red.sum", "l1tex__t_sectors_pipe_lsu_mem_global_op_atom.sum",
        "smsp__inst_executed.sum", "sm__inst_executed.avg.per_cycle_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__throughput.avg.pct_of_peak_sustained_elapsed", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread",
        "launch__grid_size", "launch__block_size", "smsp__thread_inst_executed_per_inst_executed.ratio",
        "smsp__average_warp_latency_issue_stalled_long_scoreboard.ratio", "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio", "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_membar_per_issue_active.ratio"]


def full(src, dst):
    raw = subprocess.run(["ncu", "-i", src, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units = rows[0], rows[1]
    with open(dst, "w") as f:
        f.write(f"# source: {src} (ncu --set full --clock-control none); one block per captured launch\n")
        for r in rows[2:]:
            f.write(f"\n== {r[hdr.index('Kernel Name')][:90]}\n")
            for k in KEYS:
                if k in hdr:
                    i = hdr.index(k)
                    f.write(f"  {k:82s} {r[i]:>18s} {units[i]}\n")
            if "dram__bytes_read.sum" in hdr:
                def val(k):
                    i = hdr.index(k)
                    x = float(r[i].replace(",", ""))
                    u = units[i].lower()
                    return x * (1e9 if u.startswith("gbyte") else 1e6 if u.startswith("mbyte") else 1e3 if u.startswith("kbyte") else 1)
                f.write(f"  {'traffic = dram read + write (bytes per launch)':82s} {val('dram__bytes_read.sum') + val('dram__bytes_write.sum'):18.0f}\n")
    print(open(dst).read())


def traffic(src, workload, kernel, dst):
    import json
    import os
    raw = subprocess.run(["ncu", "-i", src, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units = rows[0], rows[1]

    def val(r, k):
        i = hdr.index(k)
        x = float(r[i].replace(",", ""))
        u = units[i].lower()
        return x * (1e9 if u.startswith("gbyte") else 1e6 if u.startswith("mbyte") else 1e3 if u.startswith("kbyte") else 1)
    per = [val(r, "dram__bytes_read.sum") + val(r, "dram__bytes_write.sum") for r in rows[2:] if kernel in r[hdr.index("Kernel Name")]]
    out = json.load(open(dst)) if os.path.exists(dst) else {}
    out[workload] = {"kernel": kernel, "launches_captured": len(per), "dram_bytes_per_launch": sum(per) / len(per), "source": os.path.basename(src)}
    json.dump(out, open(dst, "w"), indent=1, sort_keys=True)
    print(out[workload])


if __name__ == "__main__":
    if sys.argv[1] == "traffic":
        traffic(*sys.argv[2:6])
    else:
        {"launches": launches, "full": full}[sys.argv[1]](sys.argv[2], sys.argv[3])
