#!/bin/bash
# box facts + the NA12878-shaped build (configs[3]) and the configs[4] query sweep on one GPU
mkdir -p gpurun_out
{ nproc; free -g; df -h /tmp /dev/shm; nvidia-smi --query-gpu=name,memory.total --format=csv; } > gpurun_out/box.txt 2>&1
KMX_TRACE=1 timeout 1500 python bench.py --workload na12878 --steps 2 --warmup 1 --no-cpu-baseline --query-sweep > gpurun_out/bench_na12878.log 2> gpurun_out/bench_na12878.err
echo "exit $?" >> gpurun_out/bench_na12878.log
tail -c 3000 gpurun_out/bench_na12878.err
tail -c 6000 gpurun_out/bench_na12878.log
