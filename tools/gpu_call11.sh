#!/bin/bash
mkdir -p gpurun_out
: > gpurun_out/r2_variants3.txt
for lib in libmezzo_b200_vH libmezzo_b200_vA libmezzo_b200_vS; do
  MZ_LIB=mezzo_b200/$lib.so timeout 600 python tools/sim_time.py sim_1m 1 2 >> gpurun_out/r2_variants3.txt 2>&1
  MZ_LIB=mezzo_b200/$lib.so timeout 300 python tools/sim_time.py sim_grid100 1 3 >> gpurun_out/r2_variants3.txt 2>&1
done
cat gpurun_out/r2_variants3.txt
