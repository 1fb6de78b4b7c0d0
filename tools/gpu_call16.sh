#!/bin/bash
# round 2: GPU suite with the transit hooks, kernel timing after the hooks went in, the driver's default bench
mkdir -p gpurun_out
(time timeout 1500 python -m pytest tests -m gpu -x -q) > gpurun_out/r2d_gputests.log 2>&1
tail -5 gpurun_out/r2d_gputests.log
timeout 300 python tools/sim_time.py sim_grid100 1 3 > gpurun_out/r2d_time.txt 2>&1
timeout 600 python tools/sim_time.py sim_1m 1 2 >> gpurun_out/r2d_time.txt 2>&1
cat gpurun_out/r2d_time.txt
(time MZ_TIMING=1 timeout 900 python bench.py) > gpurun_out/r2d_bench_default.json 2> gpurun_out/r2d_bench_default.err
tail -c 1200 gpurun_out/r2d_bench_default.json; grep "mz timing" gpurun_out/r2d_bench_default.err | tail -14; tail -4 gpurun_out/r2d_bench_default.err
