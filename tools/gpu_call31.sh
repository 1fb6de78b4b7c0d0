#!/bin/bash
# last seconds of the GPU budget: smoke() and a handful of parity cases on the library as it stands after the host-side changes
mkdir -p gpurun_out
(time timeout 10 python -c "import __graft_entry__ as g; g.smoke()") > gpurun_out/r2r_smoke.log 2>&1; tail -2 gpurun_out/r2r_smoke.log | head -1
(time timeout 25 python -m pytest tests/test_sim_gpu.py tests/test_stops.py -m gpu -x -q -k "(golden_reference and (det_ties_small or vehicle_mix_det or signals_det)) or sf_fixture_gpu") > gpurun_out/r2r_tests.log 2>&1; tail -3 gpurun_out/r2r_tests.log
