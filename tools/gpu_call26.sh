#!/bin/bash
# threaded engine build, second pass: quick parity subset + the phases of one e2e iteration at the target size
mkdir -p gpurun_out
(time timeout 120 python -m pytest tests/test_sim_gpu.py -m gpu -x -q -k "det_ties_small or congested_short or golden_reference or signals") > gpurun_out/r2n_simtests.log 2>&1
tail -3 gpurun_out/r2n_simtests.log
(time MZ_TIMING=1 timeout 600 python bench.py --workload sim_1m --no-cpu-baseline --steps 2 --warmup 1 --e2e-steps 2) > gpurun_out/r2n_bench_sim_1m.json 2> gpurun_out/r2n_bench_sim_1m.err
tail -c 300 gpurun_out/r2n_bench_sim_1m.json; grep -c timing gpurun_out/r2n_bench_sim_1m.err; tail -4 gpurun_out/r2n_bench_sim_1m.err
