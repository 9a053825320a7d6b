#!/bin/bash
mkdir -p gpurun_out
timeout 120 python tools/grid_barrier_cost.py 2>&1 | tee gpurun_out/grid_barrier_cost.log
bash tools/gpu_run40.sh
