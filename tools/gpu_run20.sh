#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q -x > gpurun_out/pytest_gpu.log 2>&1
echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
tail -4 gpurun_out/pytest_gpu.log
for cf in 0 1; do for w in rs hc14; do KMX_CLAIM_FIRST=$cf timeout 600 python bench.py --workload $w --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench_cf${cf}_$w.log 2>&1; python - <<PY
import json
l=[x for x in open('gpurun_out/bench_cf${cf}_$w.log') if x.startswith('{')]
if l:
    j=json.loads(l[-1]); print('claim_first=$cf $w value %.3g'%j['value'], 'ms/step %.2f'%j['ms_per_step'], 'insert %.2f'%j['stage_ms']['ms_insert'], j['build_stats']['insert_iterations'], j['build_stats']['insert_phase_cycles'][:7])
else: print(open('gpurun_out/bench_cf${cf}_$w.log').read()[-1500:])
PY
done; done
