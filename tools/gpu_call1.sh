#!/bin/bash
# GPU call: full GPU test-suite, then thread-count variants of the warp kernel at 1M links and on config 2
mkdir -p gpurun_out
(time timeout 1200 python -m pytest tests -m gpu -x -q) > gpurun_out/r2_gputests.log 2>&1
tail -5 gpurun_out/r2_gputests.log
: > gpurun_out/r2_variants.txt
for lib in libmezzo_b200 libmezzo_b200_t512 libmezzo_b200_t640 libmezzo_b200_t768; do
  MZ_LIB=mezzo_b200/$lib.so timeout 600 python tools/sim_time.py sim_1m 1 2 >> gpurun_out/r2_variants.txt 2>&1
  MZ_LIB=mezzo_b200/$lib.so timeout 300 python tools/sim_time.py sim_grid100 1 3 >> gpurun_out/r2_variants.txt 2>&1
done
MZ_LIB=mezzo_b200/libmezzo_b200.so timeout 600 python tools/sim_time.py sim_1m 2 2 >> gpurun_out/r2_variants.txt 2>&1
cat gpurun_out/r2_variants.txt
