#!/bin/bash
# 1 GPU, end-of-round evidence: suite, smoke, headline bench (+ reference arm), launch list + --set full capture, NA12878-shaped build + sweep
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?"; tail -3 gpurun_out/pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_ref.log 2>&1; echo "ref exit $?"; tail -c 600 gpurun_out/bench_ref.log
timeout 600 python bench.py > gpurun_out/bench.log 2> gpurun_out/bench.err; echo "bench exit $?"; tail -c 4000 gpurun_out/bench.log
CMD="python bench.py --steps 2 --warmup 1 --no-cpu-baseline"
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches.csv $CMD > gpurun_out/ncu_list.log 2>&1
echo "ncu list exit $?"
ncu --set full --clock-control none --import-source on -k regex:'insert_kernel|encode_kernel|count_kernel' -s 3 -c 3 -o gpurun_out/prof_r1_v4 -f $CMD > gpurun_out/ncu_full.log 2>&1
echo "ncu full exit $?"
ncu --set full --clock-control none --import-source on -k regex:'query_' -s 2 -c 2 -o gpurun_out/prof_r1_query_v4 -f python tools/query_only.py rs 2 > gpurun_out/ncu_query.log 2>&1
echo "ncu query exit $?"
timeout 900 python bench.py --workload hc14 --no-cpu-baseline > gpurun_out/bench_hc14.log 2> gpurun_out/bench_hc14.err; echo "hc14 exit $?"; tail -c 2500 gpurun_out/bench_hc14.log
timeout 1200 python bench.py --workload na12878 --steps 2 --warmup 1 --no-cpu-baseline --query-sweep > gpurun_out/bench_na12878.log 2> gpurun_out/bench_na12878.err
echo "na12878 exit $?"; tail -c 3500 gpurun_out/bench_na12878.log
