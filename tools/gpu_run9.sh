#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x > gpurun_out/pytest_gpu.log 2>&1
echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
tail -6 gpurun_out/pytest_gpu.log
for cfg in "20 25" "20 23" "18 25" "21 26"; do set -- $cfg; KMX_RESV_LOG2=$1 KMX_CLAIM_LOG2=$2 timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench_$1_$2.log 2>&1; python - <<PY
import json
l=[x for x in open('gpurun_out/bench_$1_$2.log') if x.startswith('{')]
if l:
    j=json.loads(l[-1]); print('resv $1 claim $2', 'value %.3g'%j['value'], 'ms/step %.2f dev %.2f'%(j['ms_per_step'], j['device_ms_per_step']), j['wall_ms_steps'], j['e2e_wall_ms_steps'], j['stage_ms'], j['build_stats']['insert_iterations'], j['build_stats']['insert_phase_cycles'], 'q %.3g %.3g'%(j['query']['value'], j['query']['e2e']['value']))
else: print(open('gpurun_out/bench_$1_$2.log').read()[-2000:])
PY
done
