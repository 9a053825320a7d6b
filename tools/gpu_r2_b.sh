#!/bin/bash
# round 2, N-GPU check of the team build: parity tests (Python ranks + in-process C++), then the strong-scaling bench
N=${1:-2}
mkdir -p gpurun_out
nvidia-smi --query-gpu=name --format=csv,noheader | wc -l
export KMX_BENCH_CACHE=/tmp/kmx_bench
timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_bench_shapes.py -m gpu -q --timeout=300 -x > gpurun_out/r2b_single_n$N.log 2>&1; echo "single exit $?"; tail -4 gpurun_out/r2b_single_n$N.log
timeout 900 python -m pytest tests/test_gpu_multi.py tests/test_gpu_cpp_api.py -m gpu -q --timeout=400 > gpurun_out/r2b_multi_n$N.log 2>&1; echo "multi exit $?"; tail -15 gpurun_out/r2b_multi_n$N.log
timeout 900 python bench.py --impl reference --steps 1 --warmup 0 > gpurun_out/r2b_bench_ref_n$N.log 2>&1; echo "ref exit $?"
for W in hc14; do
KMX_TRACE=1 timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus $N --workload $W > gpurun_out/r2b_bench_${W}_n$N.log 2> gpurun_out/r2b_bench_${W}_n$N.err; echo "$W n$N exit $?"
python - <<PY
import json
l=[x for x in open('gpurun_out/r2b_bench_${W}_n$N.log') if x.startswith('{')]
if l:
    j=json.loads(l[-1]); print('$W n$N value %.3g'%j['value'], 'ms/step %.2f dev %.2f'%(j['ms_per_step'], j['device_ms_per_step']), j['wall_ms_steps'], j['stage_ms'], j['build_stats']['insert_phase_cycles'], 'e2e %.3g %.1f ms'%(j['e2e']['value'], j['e2e']['ms_per_step']), 'q %.3g'%(j['query']['value']), j['parity'], j['roofline']['frac_of_random_sector_peak'], j['gpu_launches'])
else: print(open('gpurun_out/r2b_bench_${W}_n$N.log').read()[-2500:])
PY
grep -E "^\[kmx\]" gpurun_out/r2b_bench_${W}_n$N.err | tail -40
tail -5 gpurun_out/r2b_bench_${W}_n$N.err
done
