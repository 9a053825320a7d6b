#!/bin/bash
# round 2: suite re-check after the counter / query / diagnostics changes; read-only streaming ceiling; phase diagnostics with own-loop split
mkdir -p gpurun_out
export KMX_BENCH_CACHE=/tmp/kmx_bench
timeout 900 python -m pytest tests -m gpu -q --timeout=300 > gpurun_out/r2g_pytest_gpu.log 2>&1; echo "pytest exit $?"; tail -3 gpurun_out/r2g_pytest_gpu.log
python tools/stream_read_peak.py 2>&1 | tee gpurun_out/r2g_stream_read_peak.log
KMX_TRACE=1 timeout 900 python bench.py --no-cpu-baseline > gpurun_out/r2g_bench.log 2> gpurun_out/r2g_bench.err; echo "bench exit $?"
python - <<PY
import json
l=[x for x in open('gpurun_out/r2g_bench.log') if x.startswith('{')]
j=json.loads(l[-1]); print('hc14 value %.4g'%j['value'], 'ms/step %.2f'%j['ms_per_step'], j['stage_ms'], j['build_stats']['insert_phase_cycles'], 'e2e %.1f ms'%j['e2e']['ms_per_step'], 'q %.4g'%j['query']['value'], j['parity']['all_ranks'], j['roofline']['frac_of_random_sector_peak'], j['roofline']['traffic'])
r=j['extra']['rs']; print('rs value %.4g'%r['value'], 'ms/step %.3f'%r['ms_per_step'], r['stage_ms'], 'e2e %.2f ms'%r['e2e']['ms_per_step'], 'q %.4g'%r['query']['value'], r['parity']['all_ranks'], r['roofline']['frac_of_random_sector_peak'])
PY
grep "kmx\]" gpurun_out/r2g_bench.err | tail -12
