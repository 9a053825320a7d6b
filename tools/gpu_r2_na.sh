#!/bin/bash
# round 2: the NA12878-shaped build (3.5 G k-mers, bit_array_length > 2^32) compared with the UNMODIFIED reference, once
mkdir -p gpurun_out
export KMX_BENCH_CACHE=/tmp/kmx_bench
df -h /tmp | tail -1
( time timeout 3000 python bench.py --impl reference --workload na12878 --steps 1 --warmup 0 ) > gpurun_out/r2na_bench_ref.log 2> gpurun_out/r2na_bench_ref.err; echo "ref exit $?"; tail -c 1500 gpurun_out/r2na_bench_ref.log; tail -4 gpurun_out/r2na_bench_ref.err
KMX_BENCH_WRITE_GOLDEN=gpurun_out/r2na_golden.json timeout 1800 python bench.py --workload na12878 --steps 2 --warmup 1 --no-extra --no-cpu-baseline > gpurun_out/r2na_bench.log 2> gpurun_out/r2na_bench.err; echo "bench exit $?"; tail -c 5000 gpurun_out/r2na_bench.log; tail -5 gpurun_out/r2na_bench.err
cat gpurun_out/r2na_golden.json
