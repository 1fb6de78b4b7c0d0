#!/bin/bash
# 2 GPUs: partitioned parity tests (python/IPC and the C++ one-process seam), then the driver's N=2 bench command
mkdir -p gpurun_out
nvidia-smi -L
(time timeout 1500 python -m pytest tests/test_sim_multi_gpu.py -m gpu -q) > gpurun_out/r2e_multi_tests.log 2>&1
tail -15 gpurun_out/r2e_multi_tests.log
(time timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 2 --warmup 1) > gpurun_out/r2e_bench_n2.json 2> gpurun_out/r2e_bench_n2.err
tail -c 2500 gpurun_out/r2e_bench_n2.json; tail -5 gpurun_out/r2e_bench_n2.err
