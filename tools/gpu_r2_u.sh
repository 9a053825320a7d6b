#!/bin/bash
# round 2: the whole GPU suite on the final library (one GPU), and the smoke entry
mkdir -p gpurun_out
timeout 170 python -m pytest tests -m gpu -q -x --timeout=150 > gpurun_out/r2u_pytest_gpu.log 2>&1; echo "suite exit $?"; tail -6 gpurun_out/r2u_pytest_gpu.log
