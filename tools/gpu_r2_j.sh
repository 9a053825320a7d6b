#!/bin/bash
# round 2, 2 GPUs: team-build suite on the final code, remote-atomic microbenchmark, bench N = 2 with pread / mmap readers
mkdir -p gpurun_out
export KMX_BENCH_CACHE=/tmp/kmx_bench
timeout 900 python -m pytest tests/test_gpu_multi.py tests/test_gpu_cpp_api.py -m gpu -q --timeout=400 > gpurun_out/r2j_multi_n2.log 2>&1; echo "multi exit $?"; tail -4 gpurun_out/r2j_multi_n2.log
python tools/peer_random_peaks.py 2>&1 | tee gpurun_out/r2j_peer_random_peaks.log
for mm in 0 1; do
KMX_UPLOAD_MMAP=$mm KMX_TRACE=1 timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus 2 > gpurun_out/r2j_bench_hc14_n2_mmap$mm.log 2> gpurun_out/r2j_bench_hc14_n2_mmap$mm.err; echo "n2 mmap=$mm exit $?"
python - <<PY
import json
l=[x for x in open('gpurun_out/r2j_bench_hc14_n2_mmap$mm.log') if x.startswith('{')]
j=json.loads(l[-1]); print('hc14 n2 mmap=$mm value %.3g'%j['value'], 'ms/step %.2f dev %.2f'%(j['ms_per_step'], j['device_ms_per_step']), j['stage_ms'], j['build_stats']['insert_phase_cycles'], 'e2e %.1f ms'%j['e2e']['ms_per_step'], j['e2e']['wall_ms_steps'], 'q %.3g'%(j['query']['value']), j['parity']['all_ranks'])
PY
grep -E "upload: done" gpurun_out/r2j_bench_hc14_n2_mmap$mm.err | tail -4
done
