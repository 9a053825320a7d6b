#!/bin/bash
# 1 GPU: parity suite, then classic vs merged contested-item passes on the rs and hc14 shapes
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?"; tail -3 gpurun_out/pytest_gpu.log
for mp in 0 1; do for w in rs hc14; do
KMX_MERGED_PASSES=$mp timeout 600 python bench.py --workload $w --no-cpu-baseline --steps 4 --warmup 2 > gpurun_out/bench_${w}_mp$mp.log 2> gpurun_out/bench_${w}_mp$mp.err; echo "bench $w merged=$mp exit $?"
python - <<PY
import json
l=[x for x in open('gpurun_out/bench_${w}_mp$mp.log') if x.startswith('{')]
j=json.loads(l[-1]); print('$w merged=$mp value %.3g'%j['value'], 'ms/step %.2f'%j['ms_per_step'], 'insert %.3f'%j['stage_ms']['ms_insert'], j['build_stats']['insert_phase_cycles'], j['build_stats']['insert_iterations'])
PY
done; done
