#!/bin/bash
mkdir -p gpurun_out
(time timeout 1500 python -m pytest tests -m gpu -x -q) > gpurun_out/r2f_gputests.log 2>&1
tail -5 gpurun_out/r2f_gputests.log
(time MZ_TIMING=1 timeout 900 python bench.py) > gpurun_out/r2f_bench_default.json 2> gpurun_out/r2f_bench_default.err
tail -c 1500 gpurun_out/r2f_bench_default.json; grep "bench\] e2e" gpurun_out/r2f_bench_default.err; tail -4 gpurun_out/r2f_bench_default.err
timeout 600 python bench.py --workload sssp_grid100 --no-cpu-baseline > gpurun_out/r2f_bench_sssp_grid100.json 2> gpurun_out/r2f_bench_sssp_grid100.err
tail -c 900 gpurun_out/r2f_bench_sssp_grid100.json
