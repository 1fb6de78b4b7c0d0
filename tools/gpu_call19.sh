#!/bin/bash
mkdir -p gpurun_out
out=gpurun_out/r2g_prefetch_ab.txt
: > $out
timeout 600 python tools/sim_ab.py sim_1m 2 mezzo_b200/libmezzo_b200_nopf.so mezzo_b200/libmezzo_b200.so mezzo_b200/libmezzo_b200_pf2.so >> $out 2>&1
timeout 300 python tools/sim_ab.py sim_grid100 3 mezzo_b200/libmezzo_b200_nopf.so mezzo_b200/libmezzo_b200.so mezzo_b200/libmezzo_b200_pf2.so >> $out 2>&1
cat $out
timeout 600 python bench.py --workload sssp_cfg4 --no-cpu-baseline > gpurun_out/r2g_bench_sssp_cfg4.json 2> gpurun_out/r2g_bench_sssp_cfg4.err
tail -c 1300 gpurun_out/r2g_bench_sssp_cfg4.json; tail -2 gpurun_out/r2g_bench_sssp_cfg4.err
