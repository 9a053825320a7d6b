#!/bin/bash
# N-GPU: multi-GPU parity tests + array-owner bench (sharded Bloom inserts, peer-memory OR all-reduce)
mkdir -p gpurun_out
N=${1:-2}
timeout 600 python -m pytest tests/test_gpu_multi.py -x -q > gpurun_out/pytest_multi.log 2>&1; echo "pytest exit $?"; tail -5 gpurun_out/pytest_multi.log
for w in ${2:-hc14}; do
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus $N --steps 4 --warmup 2 --workload $w --parallelism array-owner --no-cpu-baseline > gpurun_out/bench_owner2_${w}_n$N.log 2>&1; echo "$w n$N exit $?"; python - <<PY
import json
l=[x for x in open('gpurun_out/bench_owner2_${w}_n$N.log') if x.startswith('{')]
if l:
    j=json.loads(l[-1]); print('$w n$N value %.3g'%j['value'], 'ms/step %.2f dev %.2f'%(j['ms_per_step'], j['device_ms_per_step']), j['wall_ms_steps'], j['stage_ms'], j['build_stats']['insert_phase_cycles'], 'e2e %.3g'%j['e2e']['value'], 'q %.3g'%(j['query']['value']))
else: print(open('gpurun_out/bench_owner2_${w}_n$N.log').read()[-2500:])
PY
done
