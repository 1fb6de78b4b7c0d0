import sys, time; sys.path.insert(0,'/root/repo'); sys.path.insert(0,'/root/repo/tests')
import bench, mezzo_b200
w = bench.build_sim_workload(sys.argv[1] if len(sys.argv)>1 else "sim_grid100")
sim = mezzo_b200.Simulation(w["flat"])
for cap in (1,2,4,8,16,64,1<<20):
    sim.reset(); sim.set_option(1, cap); sim.step(); s = sim.stats()
    print(f"cap {cap:8d}: {s['kernel_ms']:9.1f} ms  windows {s['n_supersteps']:7d}  us/window {1e3*s['kernel_ms']/s['n_supersteps']:6.2f}  events {s['n_events']}  trav/s {s['n_traversals']/s['kernel_ms']*1e3:.3e}", flush=True)
