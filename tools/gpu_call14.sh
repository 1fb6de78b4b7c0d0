#!/bin/bash
# lane groups (2-16 lanes per node) in the thread-per-node kernel: parity on small cases, sanitizer on a failure, then timing
mkdir -p gpurun_out
out=gpurun_out/r2_groups.txt
: > $out
for w in 8 4 2 16; do
  lib=mezzo_b200/libmezzo_b200_g$w.so
  echo "=== W=$w" >> $out
  ok=1
  for c in det_ties_small congested_short det_two_lane_congested det_gridlock; do
    MZ_LIB=$lib timeout 120 python tools/run_case.py $c 2 >> $out 2>&1 || { ok=0; echo "FAILED $c rc=$?" >> $out; break; }
  done
  if [ $ok = 0 ]; then
    MZ_LIB=$lib timeout 300 compute-sanitizer --tool memcheck --print-limit 5 python tools/run_case.py det_ties_small 2 2>&1 | grep -v "^=========     Host Frame\|^=========         in " | head -60 >> $out
    continue
  fi
  MZ_LIB=$lib timeout 900 python -m pytest tests/test_sim_gpu.py -m gpu -x -q -k thread_per_node 2>&1 | tail -4 >> $out
  MZ_LIB=$lib timeout 300 python tools/sim_time.py sim_grid100 2 2 >> $out 2>&1
  MZ_LIB=$lib timeout 600 python tools/sim_time.py sim_1m 2 1 >> $out 2>&1
done
timeout 300 python tools/sim_time.py sim_grid100 2 2 >> $out 2>&1
timeout 600 python tools/sim_time.py sim_1m 2 1 >> $out 2>&1
cat $out
