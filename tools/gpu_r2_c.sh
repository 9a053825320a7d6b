#!/bin/bash
# round 2, 8-GPU box: team-build parity at 2 / 3 / 8 ranks (+ in-process C++), strong-scaling bench at N = 8 and N = 4
mkdir -p gpurun_out
nvidia-smi --query-gpu=name --format=csv,noheader | wc -l
nvidia-smi topo -m 2>/dev/null | head -12
export KMX_BENCH_CACHE=/tmp/kmx_bench
timeout 900 python -m pytest tests/test_gpu_multi.py tests/test_gpu_cpp_api.py -m gpu -q --timeout=400 > gpurun_out/r2c_multi_n8.log 2>&1; echo "multi exit $?"; tail -15 gpurun_out/r2c_multi_n8.log
for N in 8 4; do
W=hc14
KMX_TRACE=1 timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus $N --workload $W > gpurun_out/r2c_bench_${W}_n$N.log 2> gpurun_out/r2c_bench_${W}_n$N.err; echo "$W n$N exit $?"
python - <<PY
import json
l=[x for x in open('gpurun_out/r2c_bench_${W}_n$N.log') if x.startswith('{')]
if l:
    j=json.loads(l[-1]); print('$W n$N value %.3g'%j['value'], 'ms/step %.2f dev %.2f'%(j['ms_per_step'], j['device_ms_per_step']), j['wall_ms_steps'], j['stage_ms'], j['build_stats']['insert_phase_cycles'], 'e2e %.3g %.1f ms'%(j['e2e']['value'], j['e2e']['ms_per_step']), 'q %.3g'%(j['query']['value']), j['parity'], j['roofline']['frac_of_random_sector_peak'], j['gpu_launches'])
else: print(open('gpurun_out/r2c_bench_${W}_n$N.log').read()[-2500:])
PY
grep -E "^\[kmx\]" gpurun_out/r2c_bench_${W}_n$N.err | tail -24
tail -3 gpurun_out/r2c_bench_${W}_n$N.err
done
