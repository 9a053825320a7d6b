#!/bin/bash
# round 2: KMC random access (row N4) on the GPU -- its parity tests, the accuracy report that uses it, then the whole GPU suite
mkdir -p gpurun_out
export KMX_BENCH_CACHE=/tmp/kmx_bench
timeout 300 python -m pytest tests/test_gpu_ra.py -m gpu -q --timeout=200 > gpurun_out/r2r_pytest_ra.log 2>&1; echo "ra exit $?"; tail -15 gpurun_out/r2r_pytest_ra.log
timeout 300 python tools/accuracy_report.py > gpurun_out/r2r_accuracy_rs.json 2> gpurun_out/r2r_accuracy_rs.err; echo "accuracy exit $?"; tail -5 gpurun_out/r2r_accuracy_rs.err; head -60 gpurun_out/r2r_accuracy_rs.json
timeout 300 python -m pytest tests -m gpu -q -x --timeout=200 --deselect tests/test_gpu_ra.py > gpurun_out/r2r_pytest_gpu.log 2>&1; echo "suite exit $?"; tail -3 gpurun_out/r2r_pytest_gpu.log
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
