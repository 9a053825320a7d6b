#!/bin/bash
# round 2: the random-access tests again after fixing the test database of the accuracy check
mkdir -p gpurun_out
timeout 200 python -m pytest tests/test_gpu_ra.py -m gpu -q --timeout=150 > gpurun_out/r2s_pytest_ra.log 2>&1; echo "ra exit $?"; tail -6 gpurun_out/r2s_pytest_ra.log
