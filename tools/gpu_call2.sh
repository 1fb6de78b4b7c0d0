#!/bin/bash
mkdir -p gpurun_out
(time timeout 900 python -m pytest tests/test_sim_gpu.py tests/test_files.py tests/test_abi.py -m gpu -x -q) > gpurun_out/r2_simtests.log 2>&1
tail -5 gpurun_out/r2_simtests.log
: > gpurun_out/r2_notify.txt
timeout 300 python tools/sim_time.py sim_grid100 1 3 >> gpurun_out/r2_notify.txt 2>&1
timeout 600 python tools/sim_time.py sim_1m 1 2 >> gpurun_out/r2_notify.txt 2>&1
MZ_LIB=mezzo_b200/libmezzo_b200_prof.so timeout 300 python tools/sim_profile.py sim_grid100 1 >> gpurun_out/r2_notify.txt 2>&1
MZ_LIB=mezzo_b200/libmezzo_b200_prof.so timeout 600 python tools/sim_profile.py sim_1m 1 >> gpurun_out/r2_notify.txt 2>&1
cat gpurun_out/r2_notify.txt
