#!/bin/bash
# round 2, first 1-GPU check: suite (one process per file), smoke, reference arm + headline bench on the HC14 shape
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,memory.total --format=csv,noheader | head -2
nproc
for t in test_gpu_parity test_gpu_bench_shapes test_gpu_cpp_api test_gpu_counter test_gpu_multi; do
  timeout 900 python -m pytest tests/$t.py -m gpu -q --timeout=300 > gpurun_out/r2a_$t.log 2>&1; echo "$t exit $?"; tail -4 gpurun_out/r2a_$t.log
done
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
export KMX_BENCH_CACHE=/tmp/kmx_bench
timeout 900 python bench.py --impl reference > gpurun_out/r2a_bench_ref.log 2> gpurun_out/r2a_bench_ref.err; echo "ref exit $?"; tail -c 1200 gpurun_out/r2a_bench_ref.log; tail -5 gpurun_out/r2a_bench_ref.err
timeout 1200 python bench.py > gpurun_out/r2a_bench.log 2> gpurun_out/r2a_bench.err; echo "bench exit $?"; tail -c 6000 gpurun_out/r2a_bench.log; tail -5 gpurun_out/r2a_bench.err
