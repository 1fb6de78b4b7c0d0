#!/bin/bash
mkdir -p gpurun_out
nvidia-smi -L | wc -l
(time timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 4 --steps 2 --warmup 1) > gpurun_out/r2l_bench_n4.json 2> gpurun_out/r2l_bench_n4.err
tail -c 1500 gpurun_out/r2l_bench_n4.json; tail -5 gpurun_out/r2l_bench_n4.err
