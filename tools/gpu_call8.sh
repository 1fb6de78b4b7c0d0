#!/bin/bash
mkdir -p gpurun_out
(time timeout 1200 python -m pytest tests/test_sim_gpu.py tests/test_files.py tests/test_routes.py tests/test_sssp_gpu.py tests/test_abi.py -m gpu -x -q) > gpurun_out/r2_tests8.log 2>&1
tail -4 gpurun_out/r2_tests8.log
: > gpurun_out/r2_blk.txt
timeout 300 python tools/sim_time.py sim_grid100 1 3 >> gpurun_out/r2_blk.txt 2>&1
timeout 300 python tools/sim_time.py sim_grid250 1 2 >> gpurun_out/r2_blk.txt 2>&1
timeout 600 python tools/sim_time.py sim_1m 1 2 >> gpurun_out/r2_blk.txt 2>&1
cat gpurun_out/r2_blk.txt
