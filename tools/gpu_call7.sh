#!/bin/bash
mkdir -p gpurun_out
timeout 600 ncu --section SourceCounters --section WarpStateStats --section InstructionStats --import-source on --clock-control none -k regex:k_sim_async -c 1 -o gpurun_out/r2_ncu_g250_src -f python tools/sim_time.py sim_grid250 1 1 > gpurun_out/r2_ncu_g250_src.log 2>&1
tail -2 gpurun_out/r2_ncu_g250_src.log
