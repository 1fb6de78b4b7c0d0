"""Bucket-width sweep of the batched SSSP kernel (tuning aid; run on the GPU box)."""
import sys, time
import numpy as np
sys.path.insert(0, ".")
import bench, mezzo_b200
from mezzo_b200 import _lib

name = sys.argv[1] if len(sys.argv) > 1 else "sssp_grid100"
n_trees = int(sys.argv[2]) if len(sys.argv) > 2 else 2048
w = bench.build_sssp_workload(name)
g = mezzo_b200.Graph.from_arrays(w["n_nodes"], w["n_links"], w["up"], w["dn"])
g.set_linktimes(w["T"], w["P"], w["periodlength"])
import torch
dev = torch.device("cuda", 0)
out = (torch.empty((n_trees, w["n_links"]), dtype=torch.int32, device=dev),
       torch.empty((n_trees, w["n_nodes"]), dtype=torch.int32, device=dev), None)
g.set_option(_lib.MZ_OPT_CHUNK_TREES, n_trees)
r, e, _, _ = bench.sssp_inputs(w, 0, 1, 0, n_trees)
for milli in [int(x) for x in (sys.argv[3].split(",") if len(sys.argv) > 3 else "500,1000,2000,3000,5000,8000,16000".split(","))]:
    g.set_option(_lib.MZ_OPT_DELTA_MILLI, milli)
    for rep in range(2):
        g.sssp_batch(r, e, mode=mezzo_b200.MZ_SSSP_AUTO, out=out)
    st = g.last_stats()
    print(f"delta {milli/1000:6.2f}x mean: main {st['main_kernel_ms']:9.2f} ms  all {st['kernel_ms']:9.2f} ms  steps {st['fast_steps']:7d} "
          f"eval/canon {st['evaluated_relax']/max(1,st['canonical_relax']):.3f}  redo {st['n_exact_redo']}  "
          f"{st['canonical_relax']/st['main_kernel_ms']/1e6:.2f} Grelax/s(main)  "
          f"{(21*st['canonical_relax']+12*st['reached_links'])/st['main_kernel_ms']/1e6:.0f} GB/s alg", flush=True)
