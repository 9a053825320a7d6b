#!/bin/bash
mkdir -p gpurun_out
nvidia-smi -L | wc -l
timeout 600 python -m pytest tests/test_gpu_multi.py -m gpu -q -x > gpurun_out/pytest_multi_n8.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_multi_n8.log; tail -5 gpurun_out/pytest_multi_n8.log
bash tools/gpu_run_n2c.sh 8
bash tools/gpu_run_n2c.sh 4
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29551 bench.py --gpus 8 --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench_replicas_n8.log 2>&1; echo "replicas n8 exit $?"; tail -1 gpurun_out/bench_replicas_n8.log | cut -c1-300
