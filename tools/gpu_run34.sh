#!/bin/bash
# 1 GPU: grid barrier cost, parity of the default build, A/B of block size / barrier on rs and hc14
mkdir -p gpurun_out
timeout 120 python tools/grid_barrier_cost.py 2>&1 | tee gpurun_out/grid_barrier_cost.log
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?"; tail -3 gpurun_out/pytest_gpu.log
for lib in libkmx.so libkmx_gb0.so libkmx_t512.so libkmx_t1024.so; do
for w in rs hc14; do
KMX_LIB_PATH=$PWD/kmcex_b200/$lib timeout 600 python bench.py --workload $w --no-cpu-baseline --steps 4 --warmup 2 > gpurun_out/bench_${lib}_$w.log 2> gpurun_out/bench_${lib}_$w.err; echo "bench $lib $w exit $?"
python - <<PY
import json
l=[x for x in open('gpurun_out/bench_${lib}_$w.log') if x.startswith('{')]
j=json.loads(l[-1]); print('$lib $w value %.3g'%j['value'], 'ms/step %.2f'%j['ms_per_step'], 'insert %.2f'%j['stage_ms']['ms_insert'], j['build_stats']['insert_phase_cycles'])
PY
done
done
