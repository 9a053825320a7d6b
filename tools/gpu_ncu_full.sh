#!/bin/bash
# one --set full capture of the two dominant kernels (B200_PROFILING.md recipe)
mkdir -p gpurun_out
CMD="python bench.py --steps 1 --warmup 1 --no-cpu-baseline"
timeout 600 $CMD > gpurun_out/plain.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:'insert_kernel|query_packed_kernel|encode_kernel' -c 6 -o gpurun_out/prof_r1 -f $CMD > gpurun_out/ncu_full.log 2>&1
echo "ncu exit $?"; tail -3 gpurun_out/ncu_full.log
