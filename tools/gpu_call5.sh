#!/bin/bash
mkdir -p gpurun_out
: > gpurun_out/r2_g250.txt
for lib in libmezzo_b200 libmezzo_b200_t768; do
  MZ_LIB=mezzo_b200/$lib.so timeout 300 python tools/sim_time.py sim_grid250 1 2 >> gpurun_out/r2_g250.txt 2>&1
  MZ_LIB=mezzo_b200/$lib.so timeout 300 python tools/sim_time.py sim_grid250 3 2 >> gpurun_out/r2_g250.txt 2>&1
done
cat gpurun_out/r2_g250.txt
SEC="--section SchedulerStats --section WarpStateStats --section SpeedOfLight --section MemoryWorkloadAnalysis --section InstructionStats --section LaunchStats --section Occupancy --section ComputeWorkloadAnalysis"
for lib in libmezzo_b200 libmezzo_b200_t768; do
  MZ_LIB=mezzo_b200/$lib.so timeout 500 ncu $SEC --clock-control none -k regex:k_sim_async -c 1 -o gpurun_out/r2_ncu_g250_$lib -f python tools/sim_time.py sim_grid250 1 1 > gpurun_out/r2_ncu_g250_$lib.log 2>&1
  tail -2 gpurun_out/r2_ncu_g250_$lib.log
done
