#!/bin/bash
mkdir -p gpurun_out
(time timeout 900 python -m pytest tests/test_stops.py tests/test_abi.py -m gpu -q) > gpurun_out/r2h_stops_tests.log 2>&1
tail -6 gpurun_out/r2h_stops_tests.log
(time timeout 900 python bench.py --workload sim_1m_far --steps 2 --warmup 1) > gpurun_out/r2h_bench_sim_1m_far.json 2> gpurun_out/r2h_bench_sim_1m_far.err
tail -c 1800 gpurun_out/r2h_bench_sim_1m_far.json; tail -3 gpurun_out/r2h_bench_sim_1m_far.err
