#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -q -x > gpurun_out/pytest_gpu.log 2>&1
echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
tail -3 gpurun_out/pytest_gpu.log
timeout 300 python bench.py --steps 5 --warmup 3 > gpurun_out/bench.log 2>&1; python - <<PY
import json
l=[x for x in open('gpurun_out/bench.log') if x.startswith('{')]
if l:
    j=json.loads(l[-1]); print('rs value %.3g'%j['value'], 'ms/step %.2f dev %.2f'%(j['ms_per_step'], j['device_ms_per_step']), j['stage_ms'], j['build_stats']['insert_iterations'], j['build_stats']['insert_phase_cycles'], 'e2e %.3g'%j['e2e']['value'], 'q %.3g %.3g'%(j['query']['value'], j['query']['e2e']['value']), j.get('cpu_baseline'))
else: print(open('gpurun_out/bench.log').read()[-2000:])
PY
( time timeout 1500 python bench.py --workload hc14 --steps 3 --warmup 2 --no-cpu-baseline ) > gpurun_out/bench_hc14.log 2>&1; python - <<PY
import json
l=[x for x in open('gpurun_out/bench_hc14.log') if x.startswith('{')]
if l:
    j=json.loads(l[-1]); print('hc14 n=%d value %.3g'%(j['config']['n_kmers'], j['value']), 'ms/step %.2f dev %.2f'%(j['ms_per_step'], j['device_ms_per_step']), j['stage_ms'], j['build_stats'], 'e2e', j['e2e'], j['query'])
print(open('gpurun_out/bench_hc14.log').read()[-1500:])
PY
