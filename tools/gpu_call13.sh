#!/bin/bash
# round 2 checkpoint: GPU test-suite, the driver's two bench arms exactly as the driver runs them, route generation through the seam
mkdir -p gpurun_out
(time timeout 1500 python -m pytest tests -m gpu -x -q) > gpurun_out/r2c_gputests.log 2>&1
tail -5 gpurun_out/r2c_gputests.log
(time timeout 900 python bench.py --impl reference) > gpurun_out/r2c_bench_ref.json 2> gpurun_out/r2c_bench_ref.err
tail -c 600 gpurun_out/r2c_bench_ref.json; tail -3 gpurun_out/r2c_bench_ref.err
(time timeout 900 python bench.py) > gpurun_out/r2c_bench_default.json 2> gpurun_out/r2c_bench_default.err
tail -c 1500 gpurun_out/r2c_bench_default.json; tail -3 gpurun_out/r2c_bench_default.err
(time timeout 900 python tools/route_seam.py gpurun_out/r2c_route_seam.json) > gpurun_out/r2c_route_seam.log 2>&1
tail -5 gpurun_out/r2c_route_seam.log
