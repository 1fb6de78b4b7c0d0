#!/bin/bash
# what the driver runs at round end, in its order: GPU suite, smoke, reference arm, our arm
mkdir -p gpurun_out
(time timeout 1500 python -m pytest tests -m gpu -x -q) > gpurun_out/r2k_gputests.log 2>&1
tail -4 gpurun_out/r2k_gputests.log
(time timeout 300 python -c "import __graft_entry__ as g; g.smoke()") > gpurun_out/r2k_smoke.log 2>&1
tail -3 gpurun_out/r2k_smoke.log
(time timeout 900 python bench.py --impl reference) > gpurun_out/r2k_bench_ref.json 2> gpurun_out/r2k_bench_ref.err
tail -c 400 gpurun_out/r2k_bench_ref.json; tail -3 gpurun_out/r2k_bench_ref.err
(time timeout 900 python bench.py) > gpurun_out/r2k_bench_default.json 2> gpurun_out/r2k_bench_default.err
tail -c 1000 gpurun_out/r2k_bench_default.json; tail -3 gpurun_out/r2k_bench_default.err
