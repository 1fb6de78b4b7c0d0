"""Where the end-to-end time of the link update goes (run on the GPU box; MZ_TIMING=1 prints the phases of mz_sim_create)."""
import sys, time
sys.path.insert(0, "/root/repo"); sys.path.insert(0, "/root/repo/tests")
import bench, mezzo_b200
w = bench.build_sim_workload(sys.argv[1] if len(sys.argv) > 1 else "sim_grid100")
flat = w["flat"]
for rep in range(3):
    t = [time.perf_counter()]
    sim = mezzo_b200.Simulation(flat); t.append(time.perf_counter())
    sim.step(); t.append(time.perf_counter())
    a = sim.arrivals(); lt = sim.linktimes(); t.append(time.perf_counter())
    sim.close(); t.append(time.perf_counter())
    print("create %.1f ms  run %.1f ms (kernel %.1f)  read-out %.1f ms  destroy %.1f ms" %
          tuple(1e3 * x for x in (t[1]-t[0], t[2]-t[1], sim_ms if False else 0, t[3]-t[2], t[4]-t[3])), flush=True)
