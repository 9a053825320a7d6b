#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu.log 2>&1
echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
tail -4 gpurun_out/pytest_gpu.log
for r in 2 4 8; do KMX_READERS=$r timeout 300 python bench.py --steps 6 --warmup 3 --no-cpu-baseline > gpurun_out/bench_r$r.log 2>&1; python - <<PY
import json
l=[x for x in open('gpurun_out/bench_r$r.log') if x.startswith('{')]
if l:
    j=json.loads(l[-1]); print('readers $r value %.3g'%j['value'], 'ms/step %.2f'%j['ms_per_step'], 'e2e %.3g ms %.2f'%(j['e2e']['value'], j['e2e']['ms_per_step']), j['e2e_wall_ms_steps'], 'q %.3g %.3g'%(j['query']['value'], j['query']['e2e']['value']))
else: print(open('gpurun_out/bench_r$r.log').read()[-1500:])
PY
done
