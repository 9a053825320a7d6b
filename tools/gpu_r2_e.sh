#!/bin/bash
# round 2: A/B of the L2 fetch granularity (cudaLimitMaxL2FetchGranularity) on the random-probe kernels; suite re-check
mkdir -p gpurun_out
export KMX_BENCH_CACHE=/tmp/kmx_bench
timeout 900 python -m pytest tests -m gpu -q --timeout=300 > gpurun_out/r2e_pytest_gpu.log 2>&1; echo "pytest exit $?"; tail -3 gpurun_out/r2e_pytest_gpu.log
for g in 64 32; do
echo "=== KMX_L2_FETCH=$g"
KMX_L2_FETCH=$g python tools/random_sector_peaks.py 2>&1 | grep -E "4096|512|64 MiB"
KMX_L2_FETCH=$g timeout 600 python bench.py --no-cpu-baseline --no-extra --steps 3 --warmup 2 2>/dev/null | python -c "
import json,sys
j=json.loads([x for x in sys.stdin if x.startswith('{')][-1]); print('hc14 fetch $g: ms/step %.2f'%j['ms_per_step'], j['stage_ms'], 'e2e %.1f'%j['e2e']['ms_per_step'], 'q %.4g'%j['query']['value'], 'parity', j['parity']['all_ranks'])"
KMX_L2_FETCH=$g timeout 300 python bench.py --workload rs --no-cpu-baseline --no-extra --steps 3 --warmup 2 2>/dev/null | python -c "
import json,sys
j=json.loads([x for x in sys.stdin if x.startswith('{')][-1]); print('rs fetch $g: ms/step %.3f'%j['ms_per_step'], j['stage_ms'], 'e2e %.2f'%j['e2e']['ms_per_step'], 'q %.4g'%j['query']['value'], 'parity', j['parity']['all_ranks'])"
done 2>&1 | tee gpurun_out/r2e_l2_fetch_ab.log
