#!/bin/bash
# lane-group probes (what makes several active groups per warp fail), then the round-2 ncu evidence
mkdir -p gpurun_out
out=gpurun_out/r2_groups2.txt
: > $out
export MZ_SIM_MAX_IDLE=2000000
for v in g8s g4s g8i g4i g8o g8; do
  lib=mezzo_b200/libmezzo_b200_$v.so
  for c in det_ties_small congested_short; do
    echo "=== $v $c" >> $out
    MZ_LIB=$lib timeout 60 python tools/run_case.py $c 2 2>&1 | tail -4 >> $out
    echo "rc=${PIPESTATUS[0]}" >> $out
  done
done
unset MZ_SIM_MAX_IDLE
cat $out
# ncu: launch list of the default bench command, full captures of the two dominant kernels
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02_launches_default.csv \
   python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/r02_launches_default.log 2>&1
tail -2 gpurun_out/r02_launches_default.log | cut -c1-300
timeout 900 ncu --set full --import-source on --clock-control none -k regex:k_sim_async -c 1 -f -o gpurun_out/r02_ncu_k_sim_async_sim_1m_full \
   python tools/sim_time.py sim_1m 1 1 > gpurun_out/r02_ncu_sim_full.log 2>&1
tail -2 gpurun_out/r02_ncu_sim_full.log
timeout 900 ncu --set full --import-source on --clock-control none -k regex:k_fast_run -c 1 -f -o gpurun_out/r02_ncu_k_fast_run_cfg4_full \
   python bench.py --workload sssp_cfg4 --steps 1 --warmup 0 --no-cpu-baseline > gpurun_out/r02_ncu_sssp_full.log 2>&1
tail -2 gpurun_out/r02_ncu_sssp_full.log | cut -c1-300
