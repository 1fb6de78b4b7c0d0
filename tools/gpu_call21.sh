#!/bin/bash
mkdir -p gpurun_out
out=gpurun_out/r2i_watch_ab.txt
: > $out
timeout 600 python tools/sim_ab.py sim_1m 2 mezzo_b200/libmezzo_b200_nowatch.so mezzo_b200/libmezzo_b200.so >> $out 2>&1
timeout 300 python tools/sim_ab.py sim_grid100 3 mezzo_b200/libmezzo_b200_nowatch.so mezzo_b200/libmezzo_b200.so >> $out 2>&1
cat $out
(time timeout 900 python -m pytest tests/test_sim_gpu.py -m gpu -x -q -k "warp_per_node") > gpurun_out/r2i_simtests.log 2>&1
tail -4 gpurun_out/r2i_simtests.log
