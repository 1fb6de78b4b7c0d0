"""Tuning aid (GPU box): kernel time of the link update for one workload through the library MZ_LIB names.
usage: [MZ_LIB=...] python tools/sim_time.py [workload] [kernel 0|1|2] [reps]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import bench, mezzo_b200
name = sys.argv[1] if len(sys.argv) > 1 else "sim_grid100"
w = bench.build_sim_workload(name)
sim = mezzo_b200.Simulation(w["flat"])
if len(sys.argv) > 2:
    sim.set_option(4, int(sys.argv[2]))
ms = []
for it in range(int(sys.argv[3]) if len(sys.argv) > 3 else 3):
    sim.reset(); sim.step()
    ms.append(sim.stats()["kernel_ms"])
s = sim.stats()
print(f"{os.environ.get('MZ_LIB', 'default')} {name} kernel={sys.argv[2] if len(sys.argv) > 2 else 0}: "
      f"{' '.join(f'{m:.1f}' for m in ms)} ms; {s['n_traversals']} traversals, {s['n_events']} events -> "
      f"{s['n_traversals'] / min(ms) * 1e-3:.2f} M trav/s", flush=True)
sim.close()
