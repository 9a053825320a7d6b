#!/bin/bash
# round 2, final 8-GPU evidence: team-build suite (all GPUs + in-process C++), strong-scaling bench at N = 8 / 4 / 2 (HC14), NA12878 at N = 8
mkdir -p gpurun_out
export KMX_BENCH_CACHE=/tmp/kmx_bench
timeout 900 python -m pytest tests/test_gpu_multi.py tests/test_gpu_cpp_api.py tests/test_gpu_parity.py -m gpu -q --timeout=400 -k "team_build or two_gpus or two_ranks or reference_outputs" > gpurun_out/r2l_multi_n8.log 2>&1; echo "multi exit $?"; tail -5 gpurun_out/r2l_multi_n8.log
for N in 8 4 2; do
W=hc14
timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus $N --workload $W > gpurun_out/r2l_bench_${W}_n$N.log 2> gpurun_out/r2l_bench_${W}_n$N.err; echo "$W n$N exit $?"
python - <<PY
import json
l=[x for x in open('gpurun_out/r2l_bench_${W}_n$N.log') if x.startswith('{')]
if l:
    j=json.loads(l[-1]); print('$W n$N value %.3g'%j['value'], 'ms/step %.2f dev %.2f'%(j['ms_per_step'], j['device_ms_per_step']), j['stage_ms'], j['build_stats']['insert_phase_cycles'], 'e2e %.1f ms'%j['e2e']['ms_per_step'], 'q %.3g'%(j['query']['value']), {k: j['parity'][k] for k in ('header','km.bin','rest.bin','kmer_to_occ','all_ranks','ranks_checked','replicas_equal_on_device')}, j['roofline']['frac_of_random_sector_peak'], j['gpu_launches'])
else: print(open('gpurun_out/r2l_bench_${W}_n$N.err').read()[-2500:])
PY
done
N=8; W=na12878
timeout 2400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus $N --workload $W --steps 2 --warmup 1 > gpurun_out/r2l_bench_${W}_n$N.log 2> gpurun_out/r2l_bench_${W}_n$N.err; echo "$W n$N exit $?"
python - <<PY
import json
l=[x for x in open('gpurun_out/r2l_bench_na12878_n8.log') if x.startswith('{')]
if l:
    j=json.loads(l[-1]); print('na12878 n8 value %.3g'%j['value'], 'ms/step %.1f dev %.1f'%(j['ms_per_step'], j['device_ms_per_step']), j['stage_ms'], j['build_stats']['insert_phase_cycles'], 'e2e %.1f ms'%j['e2e']['ms_per_step'], 'q %.3g'%(j['query']['value']), j['parity'])
else: print(open('gpurun_out/r2l_bench_na12878_n8.err').read()[-2500:])
PY
tail -3 gpurun_out/r2l_bench_na12878_n8.err
