#!/bin/bash
# round 2: the NA12878-shaped model built by N GPUs (team build), parity against the digests pinned in tests/golden/bench_shapes.json
N=${1:-2}
mkdir -p gpurun_out
export KMX_BENCH_CACHE=/tmp/kmx_bench
df -h /tmp | tail -1
KMX_TRACE=1 timeout 3000 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus $N --workload na12878 --steps 2 --warmup 1 > gpurun_out/r2k_bench_na12878_n$N.log 2> gpurun_out/r2k_bench_na12878_n$N.err; echo "na12878 n$N exit $?"
python - <<PY
import json
l=[x for x in open('gpurun_out/r2k_bench_na12878_n$N.log') if x.startswith('{')]
if l:
    j=json.loads(l[-1]); print('na12878 n$N value %.3g'%j['value'], 'ms/step %.1f dev %.1f'%(j['ms_per_step'], j['device_ms_per_step']), j['wall_ms_steps'], j['stage_ms'], j['build_stats']['insert_phase_cycles'], 'e2e %.1f ms'%j['e2e']['ms_per_step'], 'q %.3g'%(j['query']['value']), j['parity'])
else: print(open('gpurun_out/r2k_bench_na12878_n$N.err').read()[-3000:])
PY
grep -E "^\[kmx\]" gpurun_out/r2k_bench_na12878_n$N.err | tail -16
tail -4 gpurun_out/r2k_bench_na12878_n$N.err
