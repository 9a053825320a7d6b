#!/bin/bash
# 8 GPUs: multi-GPU parity, array-owner build on hc14 and na12878 (+ query sweep), replicas on rs
mkdir -p gpurun_out
N=8
timeout 600 python -m pytest tests/test_gpu_multi.py -x -q > gpurun_out/pytest_multi_n8.log 2>&1; echo "pytest exit $?"; tail -5 gpurun_out/pytest_multi_n8.log
run() { # name workload extra...
  name=$1; w=$2; shift 2
  timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus $N --workload $w --no-cpu-baseline "$@" > gpurun_out/bench_${name}_n$N.log 2>&1; echo "$name n$N exit $?"
  python - <<PY
import json
l=[x for x in open('gpurun_out/bench_${name}_n$N.log') if x.startswith('{')]
if l:
    j=json.loads(l[-1]); print('$name n$N value %.3g'%j['value'], 'ms/step %.2f dev %.2f'%(j['ms_per_step'], j['device_ms_per_step']), j['wall_ms_steps'], j['stage_ms'], j['build_stats']['insert_phase_cycles'], 'e2e %.3g'%j['e2e']['value'], 'q %.3g'%(j['query']['value']), j.get('query_sweep',{}).get('results'))
else: print(open('gpurun_out/bench_${name}_n$N.log').read()[-2500:])
PY
}
run owner3_hc14 hc14 --steps 4 --warmup 2 --parallelism array-owner
run replicas_rs rs --steps 5 --warmup 3
run owner3_na12878 na12878 --steps 2 --warmup 1 --parallelism array-owner --query-sweep
