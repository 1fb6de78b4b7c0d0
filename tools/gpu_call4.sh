#!/bin/bash
mkdir -p gpurun_out
timeout 1500 ncu --set full --clock-control none --import-source on -k regex:k_sim_async -c 1 -o gpurun_out/r2_ncu_sim_1m_notify -f python tools/sim_time.py sim_1m 1 1 > gpurun_out/r2_ncu_sim_1m_notify.log 2>&1
tail -3 gpurun_out/r2_ncu_sim_1m_notify.log
