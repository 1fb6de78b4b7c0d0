"""Tuning aid (GPU box): kernel time of the link update for ONE workload (built once) through several builds of the library.
usage: python tools/sim_ab.py WORKLOAD REPS lib1.so lib2.so ...   (kernel chosen by MZ_SIM_KERNEL, default 0)"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import bench
from mezzo_b200 import _lib, sim as mzsim
name, reps, libs = sys.argv[1], int(sys.argv[2]), sys.argv[3:]
w = bench.build_sim_workload(name)
for path in libs:
    _lib._lib = None
    _lib.LIB_PATH = os.path.join(ROOT, path)
    s = mzsim.Simulation(w["flat"])
    s.set_option(4, int(os.environ.get("MZ_SIM_KERNEL", "0")))
    ms = []
    for _ in range(reps):
        s.reset(); s.step()
        ms.append(s.stats()["kernel_ms"])
    st = s.stats()
    print(f"{path} {name}: {' '.join(f'{m:.1f}' for m in ms)} ms; {st['n_traversals']} traversals -> "
          f"{st['n_traversals'] / min(ms) * 1e-3:.2f} M trav/s", flush=True)
    s.close()
