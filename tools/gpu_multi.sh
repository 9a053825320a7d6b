#!/bin/bash
# N GPUs: multi-GPU parity + array-owner bench on the given workload
mkdir -p gpurun_out
N=${1:-8}; W=${2:-hc14}; TESTS=${3:-1}; EXTRA=${4:-}
if [ "$TESTS" = "1" ]; then
timeout 600 python -m pytest tests/test_gpu_multi.py -q > gpurun_out/pytest_multi_n$N.log 2>&1; echo "pytest exit $?"; tail -4 gpurun_out/pytest_multi_n$N.log
fi
timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus $N --workload $W --no-cpu-baseline --parallelism array-owner $EXTRA > gpurun_out/bench_owner4_${W}_n$N.log 2>&1; echo "$W n$N exit $?"
python - <<PY
import json
l=[x for x in open('gpurun_out/bench_owner4_${W}_n$N.log') if x.startswith('{')]
if l:
    j=json.loads(l[-1]); print('$W n$N value %.3g'%j['value'], 'ms/step %.2f dev %.2f'%(j['ms_per_step'], j['device_ms_per_step']), j['wall_ms_steps'], j['stage_ms'], j['build_stats']['insert_phase_cycles'], 'e2e %.3g'%j['e2e']['value'], 'q %.3g'%(j['query']['value']), j.get('query_sweep',{}).get('results'))
else: print(open('gpurun_out/bench_owner4_${W}_n$N.log').read()[-2500:])
PY
