#!/bin/bash
mkdir -p gpurun_out
SEC="--section SpeedOfLight --section MemoryWorkloadAnalysis --section WarpStateStats --section SchedulerStats --section SourceCounters --section InstructionStats --section LaunchStats --section Occupancy"
timeout 1100 ncu $SEC --import-source on --clock-control none -k regex:k_sim_async -c 1 -o gpurun_out/r2_ncu_sim_1m -f python tools/sim_time.py sim_1m 1 1 > gpurun_out/r2_ncu_sim_1m.log 2>&1
tail -3 gpurun_out/r2_ncu_sim_1m.log
