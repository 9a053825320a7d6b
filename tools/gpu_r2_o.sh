#!/bin/bash
# round 2: why do identical NA12878-shaped builds on one GPU differ by 10 % from step to step?  host-side timeline with the
# large allocations (address, host time, pool size) and the progress of the insert launches
mkdir -p gpurun_out
export KMX_BENCH_CACHE=/tmp/kmx_bench
KMX_TRACE=1 timeout 1000 python bench.py --workload na12878 --steps 3 --warmup 1 --no-cpu-baseline --no-extra > gpurun_out/r2o_bench_na12878_n1.log 2> gpurun_out/r2o_bench_na12878_n1.err; echo "na12878 n1 exit $?"
python - <<PY
import json
l=[x for x in open('gpurun_out/r2o_bench_na12878_n1.log') if x.startswith('{')]
if l:
    j=json.loads(l[-1]); print('na12878 n1 value %.3g'%j['value'], 'ms/step %.1f'%j['ms_per_step'], j['wall_ms_steps'], j['stage_ms'], 'e2e', j['e2e']['wall_ms_steps'])
else: print(open('gpurun_out/r2o_bench_na12878_n1.err').read()[-2500:])
PY
grep -c "kmx\]" gpurun_out/r2o_bench_na12878_n1.err
