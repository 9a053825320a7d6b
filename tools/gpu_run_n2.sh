#!/bin/bash
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/gpus.txt
timeout 600 python -m pytest tests/test_gpu_multi.py -m gpu -q -x > gpurun_out/pytest_multi.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_multi.log; tail -5 gpurun_out/pytest_multi.log
timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench_n1.log 2>&1; echo "n1 exit $?"; tail -1 gpurun_out/bench_n1.log | cut -c1-400
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29531 bench.py --gpus 2 --steps 5 --warmup 3 > gpurun_out/bench_n2.log 2>&1; echo "n2 exit $?"; tail -2 gpurun_out/bench_n2.log | cut -c1-1200
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29532 bench.py --gpus 2 --steps 1 --warmup 0 --impl reference > gpurun_out/bench_ref_n2.log 2>&1; echo "ref n2 exit $?"; tail -1 gpurun_out/bench_ref_n2.log | cut -c1-300
