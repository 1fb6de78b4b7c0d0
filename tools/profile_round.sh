#!/bin/bash
# Round profile pass (run on the GPU box): bench lines without a profiler first, then the ncu launch lists of the same
# commands and one --set full capture of each path's dominant kernel. Everything lands in gpurun_out/.
set -x
O=gpurun_out
mkdir -p $O
timeout -s KILL 900 python bench.py > $O/bench_sim_grid100.json 2> $O/bench_sim_grid100.err
timeout -s KILL 900 python bench.py --workload sssp_grid100 > $O/bench_sssp_grid100.json 2> $O/bench_sssp_grid100.err
timeout -s KILL 900 python bench.py --workload sssp_cfg4 --steps 2 --warmup 3 > $O/bench_sssp_cfg4.json 2> $O/bench_sssp_cfg4.err
timeout -s KILL 900 python bench.py --workload sssp_grid100 --sssp-mode exact --steps 2 --no-cpu-baseline > $O/bench_sssp_grid100_exact.json 2> $O/bench_sssp_grid100_exact.err
# the north star's target size on one GPU (1.02 M links, 10.2 M trips): thread-per-node kernel
timeout -s KILL 600 python bench.py --workload sim_1m --steps 2 --warmup 1 --no-cpu-baseline > $O/bench_sim_1m.json 2> $O/bench_sim_1m.err
# launch lists (per-launch times are cold-cache and serialised: only the kernels' SHARE of a step is comparable)
timeout -s KILL 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/launches_sim_grid100.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline > $O/ncu_sim_launches.log 2>&1
timeout -s KILL 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/launches_sssp_grid100.csv \
    python bench.py --workload sssp_grid100 --steps 2 --warmup 3 --no-cpu-baseline > $O/ncu_sssp_launches.log 2>&1
# full captures of the dominant kernels
timeout -s KILL 1200 ncu --set full --clock-control none --import-source on -k regex:k_sim_async --launch-skip 1 -c 1 \
    -o $O/ncu_k_sim_async python bench.py --steps 1 --warmup 3 --no-cpu-baseline > $O/ncu_sim_full.log 2>&1
timeout -s KILL 1200 ncu --set full --clock-control none --import-source on -k regex:k_fast_run --launch-skip 1 -c 1 \
    -o $O/ncu_k_fast_run python bench.py --workload sssp_grid100 --steps 1 --warmup 3 --no-cpu-baseline > $O/ncu_sssp_full.log 2>&1
ls -la $O
