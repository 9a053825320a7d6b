#!/bin/bash
mkdir -p gpurun_out
df -h /tmp | tail -1; free -g | head -2
( time timeout 1500 python bench.py --workload wgs350 --steps 2 --warmup 1 --no-cpu-baseline ) > gpurun_out/bench_wgs350.log 2>&1; python - <<PY
import json
l=[x for x in open('gpurun_out/bench_wgs350.log') if x.startswith('{')]
if l:
    j=json.loads(l[-1]); print('wgs350 n=%d value %.3g'%(j['config']['n_kmers'], j['value']), 'ms/step %.2f dev %.2f'%(j['ms_per_step'], j['device_ms_per_step']), j['stage_ms'], j['build_stats'], 'e2e', j['e2e'], j['query'], j['roofline'])
print(open('gpurun_out/bench_wgs350.log').read()[-800:])
PY
nvidia-smi --query-gpu=memory.used --format=csv
