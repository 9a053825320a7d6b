#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -x > gpurun_out/pytest_gpu.log 2>&1
echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
tail -3 gpurun_out/pytest_gpu.log
for v in libkmx.so libkmx_q3.so; do for w in rs hc14; do KMX_LIB_PATH=$PWD/kmcex_b200/$v timeout 600 python bench.py --workload $w --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench_${v}_$w.log 2>&1; python - <<PY
import json
l=[x for x in open('gpurun_out/bench_${v}_$w.log') if x.startswith('{')]
if l:
    j=json.loads(l[-1]); print('$v $w value %.3g'%j['value'], 'ms/step %.2f'%j['ms_per_step'], j['stage_ms']['ms_insert'], j['build_stats']['insert_phase_cycles'][:4], 'q %.3g %.3g'%(j['query']['value'], j['query']['e2e']['value']))
else: print(open('gpurun_out/bench_${v}_$w.log').read()[-1500:])
PY
done; done
