#!/bin/bash
# the driver's N=2 command on the threaded engine build (partitioned sim_1m + parity check across the ranks + SSSP)
mkdir -p gpurun_out
nvidia-smi -L | wc -l
(time timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus 2 --steps 2 --warmup 1) > gpurun_out/r2p_bench_n2.json 2> gpurun_out/r2p_bench_n2.err
tail -c 1200 gpurun_out/r2p_bench_n2.json; tail -5 gpurun_out/r2p_bench_n2.err
