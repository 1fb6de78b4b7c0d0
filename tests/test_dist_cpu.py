"""CPU tests of the multi-rank path (world_size 2, gloo): tree sharding + reassembly in reference loop order.
The per-rank compute is injected (the CPU oracle stands in for Graph.sssp_batch, which needs a GPU)."""
import os
import socket
import sys

import numpy as np
import pytest
import torch.multiprocessing as mp

from mezzo_b200 import dist as mzdist


def test_shard_blocks_cover_in_order():
    for n in (0, 1, 7, 8, 1000):
        for world in (1, 2, 3, 8):
            blocks = [mzdist.shard(n, r, world) for r in range(world)]
            assert blocks[0][0] == 0 and blocks[-1][1] == n
            assert all(blocks[i][1] == blocks[i + 1][0] for i in range(world - 1))
            sizes = [b[1] - b[0] for b in blocks]
            assert max(sizes) - min(sizes) <= 1


def _worker(rank, world, port, tmp):
    sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
    import torch.distributed as dist
    import cases
    import orc
    dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world)
    try:
        c = cases.make_case(7)
        spec = orc.GraphSpec(c["n_nodes"], c["n_links"], c["up"], c["dn"], c["add_order"], c["prohibitors"])
        og = orc.OracleGraph(spec)
        og.set_linktimes(c["P"], c["periodlength"], c["T"])
        ids = np.nonzero(c["up"] != -2)[0].astype(np.int32)
        roots = ids[: min(11, len(ids))]
        entries = (np.arange(len(roots)) % 3) * 100.0

        def compute(r, e):
            lp, npd, nc, _, _ = og.batch(r, e)
            return lp, npd, nc

        got = mzdist.sharded_trees(compute, roots, entries, rank, world)
        full = compute(roots, entries)
        ok = all(np.array_equal(a, b) for a, b in zip(got, full))
        open(os.path.join(tmp, f"ok{rank}"), "w").write("1" if ok else "0")
    finally:
        dist.destroy_process_group()


def test_two_rank_sharded_trees_gloo(built, tmp_path):
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    mp.spawn(_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    assert open(tmp_path / "ok0").read() == "1" and open(tmp_path / "ok1").read() == "1"


# ----------------------------------------------------------------------------------------------- partitioned link update
def test_partition_covers_and_balances():
    import simcases
    from mezzo_b200 import flatten, partition
    f = flatten.FlatNet(simcases.make("two_lane_lookback3"))
    n = f.struct.n_nodes
    for world in (1, 2, 3, 4, 8):
        r = partition.partition_nodes(n, f.link_in_node, f.link_out_node, world)
        assert r.min() == 0 and r.max() == world - 1 and len(r) == n
        cnt = np.bincount(r, minlength=world)
        assert cnt.min() > 0 and cnt.max() <= 2 * n / world + 2
        assert partition.cut_links(r, f.link_in_node, f.link_out_node).sum() < len(f.link_in_node) * (0.2 + 0.12 * world)


def native_partition(f, world):
    """mz_partition_nodes (the C++ partitioner the multi-GPU C++ seam uses; pure host code, no device needed)"""
    import ctypes as C
    import mezzo_b200
    nr = np.empty(f.struct.n_nodes, np.int32)
    rc = mezzo_b200.lib().mz_partition_nodes(C.byref(f.struct), world, nr.ctypes.data_as(C.c_void_p))
    assert rc == 0, mezzo_b200.lib().mz_last_error()
    return nr


def test_native_partition_covers_and_balances(built):
    import simcases
    from mezzo_b200 import flatten, partition
    f = flatten.FlatNet(simcases.make("two_lane_lookback3"))
    n = f.struct.n_nodes
    for world in (1, 2, 3, 4, 8):
        r = native_partition(f, world)
        assert r.min() == 0 and r.max() == world - 1 and len(r) == n
        assert np.bincount(r, minlength=world).min() > 0
        assert partition.cut_links(r, f.link_in_node, f.link_out_node).sum() < len(f.link_in_node) * (0.2 + 0.12 * world)


def _sim_worker(rank, world, port, tmp, name, native=False):
    sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
    import torch.distributed as dist
    import orc
    import simcases
    from mezzo_b200 import flatten
    dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world)
    try:
        f = flatten.FlatNet(simcases.make(name))
        ps = mzdist.PartitionedSimulation(f, rank, world, orc.EmuPartBackend(f"/mzemu_{port}_"),
                                          node_rank=native_partition(f, world) if native else None)
        end = ps.step()
        res = dict(end=end, stats=ps.stats(), arrivals=ps.arrivals(), linktimes=ps.linktimes(), exits=ps.exit_log())
        ps.close()
        if rank == 0:
            o = orc.oracle_sim(f)
            oe = np.sort(o["exits"], order=["vid", "entry"])
            ok = len(oe) == len(res["exits"]) and all(np.array_equal(oe[k], res["exits"][k]) for k in ("vid", "link", "entry", "exit"))
            ok = ok and np.array_equal(o["arrivals"], res["arrivals"]) and np.array_equal(o["linktimes"], res["linktimes"])
            import simcases as _sc
            keys = ("n_traversals", "n_exits", "n_arrived", "end_time") + (() if name in _sc.SIGNALS else ("n_events",))
            ok = ok and all(o["stats"][k] == res["stats"][k] for k in keys)
            open(os.path.join(tmp, "ok"), "w").write("1" if ok else "0")
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("name,world,native", [("det_ties_small", 2, False), ("congested_short", 2, False),
                                               ("two_lane_lookback3", 3, False), ("signals_det", 2, False),
                                               ("rate_changes_det", 3, False), ("det_ties_small", 3, True),
                                               ("det_two_lane_congested", 8, True)])
def test_partitioned_link_update_gloo(built, tmp_path, name, world, native):
    """The multi-rank link update (node partition, arenas shared between ranks, pairwise key synchronisation, merged
    outputs, the single event beyond the horizon) on CPU processes under gloo must be bit-identical to the sequential
    oracle — including the tie-heavy network where event order across the cut matters."""
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    mp.spawn(_sim_worker, args=(world, port, str(tmp_path), name, native), nprocs=world, join=True)
    assert open(tmp_path / "ok").read() == "1"
