#!/bin/bash
# round 2: C++ boundary of the random-access row (include/kmc_ra.hpp) + the random-access tests on the final library
mkdir -p gpurun_out
timeout 200 python -m pytest tests/test_gpu_cpp_api.py tests/test_gpu_ra.py -m gpu -q --timeout=150 > gpurun_out/r2t_pytest_ra_cpp.log 2>&1; echo "exit $?"; tail -12 gpurun_out/r2t_pytest_ra_cpp.log
