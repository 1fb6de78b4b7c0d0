#!/bin/bash
mkdir -p gpurun_out
(time timeout 900 python -m pytest tests/test_sssp_gpu.py tests/test_routes.py -m gpu -x -q) > gpurun_out/r2_sssptests.log 2>&1
tail -4 gpurun_out/r2_sssptests.log
(time timeout 900 python -m pytest tests/test_fullsize_gpu.py -m gpu -x -q -k "config4 or config3") >> gpurun_out/r2_sssptests.log 2>&1
tail -4 gpurun_out/r2_sssptests.log
timeout 600 python bench.py --workload sssp_grid100 --steps 3 --warmup 2 --no-cpu-baseline > gpurun_out/r2_bench_sssp_grid100.json 2> gpurun_out/r2_bench_sssp_grid100.err
timeout 900 python bench.py --workload sssp_cfg4 --steps 3 --warmup 2 --no-cpu-baseline > gpurun_out/r2_bench_sssp_cfg4.json 2> gpurun_out/r2_bench_sssp_cfg4.err
python - <<'P'
import json
for n in ("sssp_grid100","sssp_cfg4"):
    try:
        d=json.loads(open(f"gpurun_out/r2_bench_{n}.json").read().strip().splitlines()[-1])
        print(n, "value %.3g ms %.1f e2e %.3g frac %.3f main_ms %.1f all_ms %.1f steps %s evalratio %.3f redo %s" % (d["value"], d["ms_per_step"], d["e2e"]["value"], d["roofline"]["frac"], d["roofline"]["kernel_ms_per_launch"], d["roofline"].get("all_kernels_ms_per_step",0), d["work"].get("near_far_steps_per_step"), d["roofline"].get("evaluated_over_canonical",0), d["work"].get("exact_redo_trees_per_step")))
    except Exception as e:
        print(n, "failed", e); print(open(f"gpurun_out/r2_bench_{n}.err").read()[-2000:])
P
