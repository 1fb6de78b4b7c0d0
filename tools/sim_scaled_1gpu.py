"""How long does the N-GPU weak-scaling network of bench.py take on ONE GPU? (reference point for the multi-GPU runs)"""
import sys; sys.path.insert(0, "/root/repo"); sys.path.insert(0, "/root/repo/tests")
import bench, mezzo_b200
for world in [int(x) for x in sys.argv[1].split(",")]:
    w = bench.build_sim_workload("sim_grid100", world=world)
    sim = mezzo_b200.Simulation(w["flat"])
    for rep in range(2):
        sim.reset(); sim.step(); s = sim.stats()
    print(f"network for {world} GPUs on one GPU: links {w['flat'].struct.n_links} nodes {w['flat'].struct.n_nodes} "
          f"{s['kernel_ms']:.1f} ms  {s['n_traversals']/s['kernel_ms']*1e3:.4g} trav/s", flush=True)
    sim.close()
