#!/bin/bash
# round 2: compute-sanitizer over a small slice of the parity suite (memcheck, then racecheck + synccheck on the shared-memory
# kernels), then the NA12878 shape on one GPU with the host-side timeline (rest stage after freeing the item stream first)
mkdir -p gpurun_out
export KMX_BENCH_CACHE=/tmp/kmx_bench
SAN=/usr/local/cuda/bin/compute-sanitizer
SEL="tiny_ci1 or empty_and_tiny or fresh_seed or error_corners"
timeout 600 $SAN --tool memcheck --error-exitcode 9 --print-limit 20 python -m pytest tests/test_gpu_parity.py -m gpu -q -x --timeout=500 -k "$SEL" > gpurun_out/r2n_memcheck_parity.log 2>&1; echo "memcheck parity exit $?"; tail -4 gpurun_out/r2n_memcheck_parity.log
timeout 400 $SAN --tool memcheck --error-exitcode 9 --print-limit 20 python -m pytest tests/test_gpu_counter.py -m gpu -q -x --timeout=300 -k "31-1-1023 or gzip" > gpurun_out/r2n_memcheck_counter.log 2>&1; echo "memcheck counter exit $?"; tail -4 gpurun_out/r2n_memcheck_counter.log
timeout 500 $SAN --tool racecheck --error-exitcode 9 --print-limit 20 python -m pytest tests/test_gpu_parity.py -m gpu -q -x --timeout=400 -k "tiny_ci1 and (listing or byte_identical or ascii)" > gpurun_out/r2n_racecheck_parity.log 2>&1; echo "racecheck parity exit $?"; tail -4 gpurun_out/r2n_racecheck_parity.log
timeout 400 $SAN --tool synccheck --error-exitcode 9 --print-limit 20 python -m pytest tests/test_gpu_parity.py -m gpu -q -x --timeout=300 -k "tiny_ci1 and (listing or byte_identical or ascii)" > gpurun_out/r2n_synccheck_parity.log 2>&1; echo "synccheck parity exit $?"; tail -4 gpurun_out/r2n_synccheck_parity.log
KMX_TRACE=1 timeout 1200 python bench.py --workload na12878 --steps 2 --warmup 1 --no-cpu-baseline --no-extra > gpurun_out/r2n_bench_na12878_n1.log 2> gpurun_out/r2n_bench_na12878_n1.err; echo "na12878 n1 exit $?"
python - <<PY
import json
l=[x for x in open('gpurun_out/r2n_bench_na12878_n1.log') if x.startswith('{')]
if l:
    j=json.loads(l[-1]); print('na12878 n1 value %.3g'%j['value'], 'ms/step %.1f dev %.1f'%(j['ms_per_step'], j['device_ms_per_step']), j['stage_ms'], 'e2e %.1f ms'%j['e2e']['ms_per_step'], 'q %.3g'%(j['query']['value']), j['parity'])
else: print(open('gpurun_out/r2n_bench_na12878_n1.err').read()[-2500:])
PY
grep "kmx\]" gpurun_out/r2n_bench_na12878_n1.err | tail -24
