#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x > gpurun_out/pytest_gpu.log 2>&1
echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
tail -4 gpurun_out/pytest_gpu.log
for w in rs hc14; do
timeout 900 python bench.py --workload $w --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench_$w.log 2>&1; python - <<PY
import json
l=[x for x in open('gpurun_out/bench_$w.log') if x.startswith('{')]
if l:
    j=json.loads(l[-1]); print('$w value %.3g'%j['value'], 'ms/step %.2f dev %.2f'%(j['ms_per_step'], j['device_ms_per_step']), j['stage_ms'], j['build_stats']['insert_iterations'], j['build_stats']['insert_phase_cycles'], 'e2e %.3g'%j['e2e']['value'], 'q %.3g %.3g'%(j['query']['value'], j['query']['e2e']['value']), j['roofline']['frac'], j['roofline']['frac_of_random_sector_peak'])
else: print(open('gpurun_out/bench_$w.log').read()[-2000:])
PY
done
