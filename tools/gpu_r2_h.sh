#!/bin/bash
# round 2: counting pass after the consume-side rewrite: suite, stage times, ncu DRAM throughput of count_kernel
mkdir -p gpurun_out
export KMX_BENCH_CACHE=/tmp/kmx_bench
timeout 900 python -m pytest tests -m gpu -q --timeout=300 > gpurun_out/r2h_pytest_gpu.log 2>&1; echo "pytest exit $?"; tail -3 gpurun_out/r2h_pytest_gpu.log
timeout 900 python bench.py --no-cpu-baseline --steps 3 --warmup 2 > gpurun_out/r2h_bench.log 2> gpurun_out/r2h_bench.err; echo "bench exit $?"
python - <<PY
import json
l=[x for x in open('gpurun_out/r2h_bench.log') if x.startswith('{')]
j=json.loads(l[-1]); print('hc14 ms/step %.2f'%j['ms_per_step'], j['stage_ms'], 'e2e %.1f ms'%j['e2e']['ms_per_step'], j['parity']['all_ranks'])
r=j['extra']['rs']; print('rs ms/step %.3f'%r['ms_per_step'], r['stage_ms'], 'e2e %.2f ms'%r['e2e']['ms_per_step'], r['parity']['all_ranks'])
PY
timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed,sm__inst_executed.avg.per_cycle_elapsed,smsp__inst_executed.sum --clock-control none -k regex:'count_kernel' -s 1 -c 1 --csv --log-file gpurun_out/r2h_ncu_count.csv python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-extra > /dev/null 2>&1
grep -E "count_kernel" gpurun_out/r2h_ncu_count.csv | awk -F'","' '{print $(NF-2), $(NF-1), $NF}'
