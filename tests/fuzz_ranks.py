"""TEST INFRASTRUCTURE (CPU, not collected by pytest): randomised parity sweep of the PARTITIONED link update.
Random small networks with random feature mixes, cut into 2-5 ranks by the native partitioner or at random; every rank is a
process running the engine's own code compiled for the host (tests/emu), arenas in POSIX shared memory, orchestration
mezzo_b200/dist.py under gloo — the merged result must equal the sequential C oracle bit for bit.
usage: python tests/fuzz_ranks.py [seed] [count]"""
import os
import socket
import sys
import tempfile

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
sys.path.insert(0, HERE)
import numpy as np


def worker(rank, world, port, tmp, kw, mode, pseed):
    import torch.distributed as dist
    import orc
    from mezzo_b200 import dist as mzdist, flatten, synth
    from test_dist_cpu import native_partition
    dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world)
    try:
        f = flatten.FlatNet(synth.sim_grid(**kw))
        if mode == "native":
            node_rank = native_partition(f, world)
        elif mode == "random":  # the protocol must hold for ANY assignment, however bad the cut
            node_rank = np.random.default_rng(pseed).integers(0, world, f.struct.n_nodes).astype(np.int32)
        else:
            node_rank = None  # mezzo_b200/partition.py
        ps = mzdist.PartitionedSimulation(f, rank, world, orc.EmuPartBackend(f"/mzfz_{port}_"), node_rank=node_rank)
        ps.step()
        res = dict(stats=ps.stats(), arrivals=ps.arrivals(), linktimes=ps.linktimes(), exits=ps.exit_log())
        ps.close()
        if rank == 0:
            o = orc.oracle_sim(f)
            oe = np.sort(o["exits"], order=["vid", "entry"])
            ok = len(oe) == len(res["exits"]) and all(np.array_equal(oe[k], res["exits"][k]) for k in ("vid", "link", "entry", "exit"))
            ok = ok and np.array_equal(o["arrivals"], res["arrivals"]) and np.array_equal(o["linktimes"], res["linktimes"])
            ok = ok and all(o["stats"][k] == res["stats"][k] for k in ("n_traversals", "n_exits", "n_arrived", "end_time"))
            open(os.path.join(tmp, "ok"), "w").write(f"{int(ok)} {len(oe)}")
    finally:
        dist.destroy_process_group()


if __name__ == "__main__":
    import torch.multiprocessing as mp
    rng = np.random.default_rng(int(sys.argv[1]) if len(sys.argv) > 1 else 0)
    N = int(sys.argv[2]) if len(sys.argv) > 2 else 10
    bad = 0
    for it in range(N):
        det = bool(rng.integers(2))
        kw = dict(nx=int(rng.integers(4, 8)), ny=int(rng.integers(4, 8)), seed=int(rng.integers(1000)), od_every=2,
                  n_od=int(rng.integers(10, 40)), trips_per_hour=int(rng.integers(200, 2500)), horizon=float(rng.choice([600.0, 1200.0])),
                  two_lane_prob=float(rng.choice([0.0, 0.3, 0.6])), lookback=int(rng.choice([2, 4, 20])),
                  len_lo=int(rng.choice([40, 60, 150, 300])), connector_len=int(rng.choice([30, 50, 150, 250])),
                  signals=int(rng.choice([0, 0, 4, 10])), rate_changes=int(rng.choice([0, 0, 20])),
                  incidents=int(rng.choice([0, 0, 2])), demand_slices=int(rng.choice([0, 0, 2])),
                  alt_routes=int(rng.choice([1, 1, 3])), n_periods=int(rng.choice([2, 4])))
        kw["len_hi"] = kw["len_lo"] + int(rng.choice([0, 50, 300]))
        if rng.integers(3) == 0:
            kw["vehicle_classes"] = [(float(p), float(ln)) for p, ln in zip(rng.uniform(0.05, 0.5, 3), rng.choice([4.0, 6.0, 9.5, 12.0, 18.0], 3, replace=False))]
        if det:
            kw.update(server_type=2, server_mu=float(rng.choice([1.0, 1.5, 2.0])), server_sd=0.0, dest_server=(2, float(rng.choice([0.3, 0.5, 1.0])), 0.0, 0.0))
        world = int(rng.integers(2, 6))
        mode = str(rng.choice(["native", "python", "random"]))
        s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
        tmp = tempfile.mkdtemp()
        try:
            mp.spawn(worker, args=(world, port, tmp, kw, mode, int(rng.integers(1 << 30))), nprocs=world, join=True)
            ok, n = open(os.path.join(tmp, "ok")).read().split()
        except Exception as e:  # noqa: BLE001
            ok, n = "0", f"EXC {type(e).__name__}: {str(e)[-1500:]}"
        if ok != "1":
            bad += 1
        print(it, "OK " if ok == "1" else "MISMATCH", n, "exits", world, "ranks", mode, {k: kw[k] for k in ("signals", "rate_changes", "incidents", "demand_slices", "trips_per_hour")},
              "det" if det else "stoch", "mix" if "vehicle_classes" in kw else "", flush=True)
    print("bad", bad)
