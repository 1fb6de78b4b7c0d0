#!/bin/bash
# final validation, part 1: the GPU suite and smoke() as the driver runs them
mkdir -p gpurun_out
(time timeout 1500 python -m pytest tests -m gpu -q) > gpurun_out/r2q_gputests.log 2>&1
tail -6 gpurun_out/r2q_gputests.log
(time timeout 300 python -c "import __graft_entry__ as g; g.smoke()") > gpurun_out/r2q_smoke.log 2>&1
tail -4 gpurun_out/r2q_smoke.log
