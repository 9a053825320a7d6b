#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x > gpurun_out/pytest_gpu.log 2>&1
echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
tail -4 gpurun_out/pytest_gpu.log
for lg in 20 21 22; do KMX_RESV_LOG2=$lg timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench_resv$lg.log 2>&1; python - <<PY
import json
l=[x for x in open('gpurun_out/bench_resv$lg.log') if x.startswith('{')]
if l:
    j=json.loads(l[-1]); print('resv$lg', 'value %.3g'%j['value'], 'ms/step %.2f dev %.2f'%(j['ms_per_step'], j['device_ms_per_step']), j['stage_ms'], j['build_stats'], 'e2e', j['e2e'], j['query'])
else: print(open('gpurun_out/bench_resv$lg.log').read()[-2000:])
PY
done
CMD="python bench.py --steps 1 --warmup 1 --no-cpu-baseline"
timeout 600 $CMD > gpurun_out/plain.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:'query_packed_kernel' -c 2 -o gpurun_out/prof_r1_query -f $CMD > gpurun_out/ncu_full_q.log 2>&1
echo "ncu exit $?"
