#!/bin/bash
mkdir -p gpurun_out
: > gpurun_out/r2_notify2.txt
for lib in libmezzo_b200 libmezzo_b200_t512 libmezzo_b200_t768; do
  MZ_LIB=mezzo_b200/$lib.so timeout 600 python tools/sim_time.py sim_1m 1 2 >> gpurun_out/r2_notify2.txt 2>&1
done
for lib in libmezzo_b200_prof libmezzo_b200_pt768; do
MZ_LIB=mezzo_b200/$lib.so timeout 600 python tools/sim_profile.py sim_1m 1 >> gpurun_out/r2_notify2.txt 2>&1
done
MZ_LIB=mezzo_b200/libmezzo_b200_t768.so timeout 300 python tools/sim_time.py sim_grid100 1 3 >> gpurun_out/r2_notify2.txt 2>&1
cat gpurun_out/r2_notify2.txt
