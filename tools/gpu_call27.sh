#!/bin/bash
# threaded engine build, third pass: phases of one e2e iteration at the target size; then a --set full capture of the
# node-visit kernel as it stands at the end of round 2 (config 2: one launch = one simulated hour, ~205 ms)
mkdir -p gpurun_out
(time MZ_TIMING=1 timeout 600 python bench.py --workload sim_1m --no-cpu-baseline --steps 2 --warmup 1 --e2e-steps 2) > gpurun_out/r2o_bench_sim_1m.json 2> gpurun_out/r2o_bench_sim_1m.err
tail -c 300 gpurun_out/r2o_bench_sim_1m.json; tail -4 gpurun_out/r2o_bench_sim_1m.err
(time timeout 200 ncu --set full --import-source on --clock-control none -k regex:k_sim_async -c 1 -f -o gpurun_out/r02_ncu_k_sim_async_cfg2_full \
   python tools/sim_time.py sim_grid100 1 1) > gpurun_out/r02_ncu_sim_cfg2_full.log 2>&1
tail -5 gpurun_out/r02_ncu_sim_cfg2_full.log
ls -la gpurun_out/r02_ncu_k_sim_async_cfg2_full.ncu-rep
