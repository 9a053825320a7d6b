#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x > gpurun_out/pytest_gpu.log 2>&1
echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
tail -4 gpurun_out/pytest_gpu.log
KMX_TRACE=1 timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench.log 2> gpurun_out/bench_trace.log; python - <<PY
import json
l=[x for x in open('gpurun_out/bench.log') if x.startswith('{')]
if l:
    j=json.loads(l[-1]); print('value %.3g'%j['value'], 'ms/step %.2f dev %.2f'%(j['ms_per_step'], j['device_ms_per_step']), j['wall_ms_steps'], j['e2e_wall_ms_steps'], j['stage_ms'], 'e2e', j['e2e'], j['query'])
else: print(open('gpurun_out/bench.log').read()[-2000:])
PY
tail -60 gpurun_out/bench_trace.log
