"""Tuning aid (run on the GPU box): cycle breakdown of the link-update kernel from the profile build
(make -C mezzo_b200/csrc prof; MZ_LIB=mezzo_b200/libmezzo_b200_prof.so python tools/sim_profile.py [workload] [kernel])."""
import ctypes as C, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import bench, mezzo_b200
w = bench.build_sim_workload(sys.argv[1] if len(sys.argv) > 1 else "sim_grid100")
sim = mezzo_b200.Simulation(w["flat"])
if len(sys.argv) > 2:
    sim.set_option(4, int(sys.argv[2]))
for it in range(2):
    sim.reset(); sim.step()
s = sim.stats()
ev = (C.c_uint64 * 56)()
sim.l.mz_sim_debug_prof_events.argtypes = [C.c_void_p, C.POINTER(C.c_uint64 * 56)]
sim.l.mz_sim_debug_prof_events(sim._h, C.byref(ev))
print(f"{s['kernel_ms']:.1f} ms, {s['n_events']} events, {s['n_supersteps']} visits, {s['n_traversals']} traversals")
for i, nm in enumerate(["turn fail", "turn ok", "sink fail", "sink ok", "origin"]):
    if ev[3 * i + 1]:
        print(f"  event {nm:10s} n {ev[3*i+1]:10d}  mean {ev[3*i]/ev[3*i+1]:8.0f} cyc  max {ev[3*i+2]:8d}")
for c, nm in enumerate(["interior (class 0)", "region boundary (class 1)", "rank boundary (class 2)"]):
    n = ev[24 + 4 * c]
    if n:
        print(f"  visits of {nm:26s} n {n:10d}  before loop {ev[25+4*c]/n:7.0f}  event loop {ev[26+4*c]/n:7.0f}  publish {ev[27+4*c]/n:7.0f} cyc")
if ev[44]:
    alive, inv, sweeps, calls, warps = ev[40], ev[41], ev[42], ev[43], ev[44]
    print(f"  warps {warps}: alive {alive/warps:.3e} cyc each, inside async_visit {100*inv/alive:.1f} %, sweeps {sweeps/warps:.0f} per warp "
          f"({(alive-inv)/max(1,sweeps):.0f} cyc per sweep outside visits), visit calls {calls} ({calls/max(1,s['n_supersteps']):.2f} per executed visit), "
          f"idle iterations {ev[45]/warps:.0f} per warp")
sim.close()
