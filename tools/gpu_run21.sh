#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu.log 2>&1
echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
tail -4 gpurun_out/pytest_gpu.log
timeout 600 python bench.py --steps 5 --warmup 3 > gpurun_out/bench.log 2>&1; python - <<PY
import json
l=[x for x in open('gpurun_out/bench.log') if x.startswith('{')]
if l:
    j=json.loads(l[-1]); print('value %.3g'%j['value'], 'ms/step %.2f'%j['ms_per_step'], 'e2e %.3g'%j['e2e']['value'], j['query'], j['roofline']['frac'], j['roofline']['frac_of_random_sector_peak'], j['gpu_launches'], j['cpu_baseline'])
else: print(open('gpurun_out/bench.log').read()[-1500:])
PY
timeout 300 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_ref.log 2>&1; tail -1 gpurun_out/bench_ref.log | cut -c1-400
