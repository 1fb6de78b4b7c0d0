"""Debug aid (GPU box): run one simcases network through a chosen node-visit kernel and compare with the CPU oracle.
usage: python tools/run_case.py CASE [kernel 0|1|2]"""
import sys
sys.path.insert(0, "/root/repo"); sys.path.insert(0, "/root/repo/tests")
import numpy as np
import mezzo_b200, orc, simcases
from mezzo_b200 import flatten
name = sys.argv[1]
kernel = int(sys.argv[2]) if len(sys.argv) > 2 else 0
f = flatten.FlatNet(simcases.make(name))
o = orc.oracle_sim(f)
sim = mezzo_b200.Simulation(f)
sim.set_option(mezzo_b200.MZ_SIM_OPT_KERNEL, kernel)
sim.step()
ex, oe = sim.exit_log(), np.sort(o["exits"], order=["vid", "entry"])
st = sim.stats()
print({k: (st[k], o["stats"][k]) for k in ("n_traversals", "n_exits", "n_arrived", "n_events", "end_time")})
same = len(ex) == len(oe) and all(np.array_equal(ex[k], oe[k]) for k in ("vid", "link"))
print("exit log: same sequence", same, "max |dt|", float(np.abs(ex["exit"] - oe["exit"]).max()) if same else None)
sim.close()
exact = same and np.array_equal(ex["exit"], oe["exit"]) and np.array_equal(ex["entry"], oe["entry"])
print("bit-identical exit log:", exact)
sys.exit(0 if same and all(st[k] == o["stats"][k] for k in ("n_traversals", "n_exits", "n_arrived", "end_time")) else 1)
