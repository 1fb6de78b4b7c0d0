#!/bin/bash
# final validation, part 2: the driver's default bench command
mkdir -p gpurun_out
(time timeout 160 python bench.py) > gpurun_out/r2q_bench_default.json 2> gpurun_out/r2q_bench_default.err
tail -c 600 gpurun_out/r2q_bench_default.json; tail -4 gpurun_out/r2q_bench_default.err
