"""Tuning aid (run on the GPU box): per-window cost of the link update for different event caps / block counts."""
import ctypes as C, sys, time; sys.path.insert(0,'/root/repo'); sys.path.insert(0,'/root/repo/tests')
import bench, mezzo_b200
w = bench.build_sim_workload(sys.argv[1] if len(sys.argv)>1 else "sim_grid100")
caps = [int(x) for x in (sys.argv[2].split(",") if len(sys.argv)>2 else "1,2,4".split(","))]
blocks = [int(x) for x in (sys.argv[3].split(",") if len(sys.argv)>3 else "0".split(","))]
sim = mezzo_b200.Simulation(w["flat"])
sim.l.mz_sim_debug_prof.argtypes = [C.c_void_p, C.POINTER(C.c_uint64 * 3)]
for nb in blocks:
  for cap in caps:
    sim.reset(); sim.set_option(1, cap); sim.set_option(3, nb); sim.step(); s = sim.stats()
    pr = (C.c_uint64 * 3)(); sim.l.mz_sim_debug_prof(sim._h, C.byref(pr)); nw = s['n_supersteps']
    print(f"blocks {nb:4d} cap {cap:8d}: {s['kernel_ms']:9.1f} ms  windows {nw:7d}  us/window {1e3*s['kernel_ms']/nw:6.2f}  "
          f"cycles/window phase1 {pr[0]/nw:7.0f} phase2 {pr[1]/nw:7.0f} barrier {pr[2]/nw:7.0f}  trav/s {s['n_traversals']/s['kernel_ms']*1e3:.3e}", flush=True)
    ev = (C.c_uint64 * 24)(); sim.l.mz_sim_debug_prof_events.argtypes = [C.c_void_p, C.POINTER(C.c_uint64 * 24)]
    sim.l.mz_sim_debug_prof_events(sim._h, C.byref(ev))
    for i, nm in enumerate(["turn fail", "turn ok", "sink fail", "sink ok", "origin"]):
        if ev[3*i+1]: print(f"      {nm:10s} n {ev[3*i+1]:9d}  mean {ev[3*i]/ev[3*i+1]:8.0f} cyc  max {ev[3*i+2]:8d} cyc")
