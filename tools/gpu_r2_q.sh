#!/bin/bash
# round 2, 2 GPUs: pre-flight of the team path after the single-GPU survivor-list change (team tests, in-process C++, bench N = 2)
mkdir -p gpurun_out
export KMX_BENCH_CACHE=/tmp/kmx_bench
timeout 600 python -m pytest tests/test_gpu_multi.py tests/test_gpu_cpp_api.py -m gpu -q --timeout=300 > gpurun_out/r2q_multi.log 2>&1; echo "multi exit $?"; tail -3 gpurun_out/r2q_multi.log
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus 2 --steps 3 --warmup 2 > gpurun_out/r2q_bench_hc14_n2.log 2> gpurun_out/r2q_bench_hc14_n2.err; echo "n2 exit $?"
python - <<PY
import json
l=[x for x in open('gpurun_out/r2q_bench_hc14_n2.log') if x.startswith('{')]
if l:
    j=json.loads(l[-1]); print('hc14 n2 ms/step %.2f'%j['ms_per_step'], j['stage_ms'], 'e2e %.1f ms'%j['e2e']['ms_per_step'], j['parity'])
else: print(open('gpurun_out/r2q_bench_hc14_n2.err').read()[-2500:])
PY
