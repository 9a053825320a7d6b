#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_multi.py -m gpu -q -x > gpurun_out/pytest_multi.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_multi.log; tail -30 gpurun_out/pytest_multi.log
timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "byte_identical or fresh" > gpurun_out/pytest_single.log 2>&1; tail -3 gpurun_out/pytest_single.log
