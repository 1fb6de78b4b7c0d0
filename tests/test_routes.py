"""Route generation end to end (hot path 2 at the system level, SURVEY.md §8 a16 + f2): the host mirror of
Network::shortest_paths_all (mezzo_b200/routes.py) against the routes.dat the UNMODIFIED reference rewrites after a
calc_paths=1 run, and the simulation that then runs on those route sets with route choice.
CPU tests inject the oracle as the tree engine; the GPU test drives mz_sssp_routes."""
import os
import re

import numpy as np
import pytest

import orc
import refrun
from mezzo_b200 import flatten, routes as mzroutes, synth

CASES = {
    # no routes given at all: everything comes from the SSSP (3 reruns: entry times 0, 900, 1800 s), time-varying link times
    "all_from_sssp": dict(nx=7, ny=6, seed=31, od_every=2, n_od=30, trips_per_hour=400, horizon=3600.0, keep_routes=0.0,
                          hist_variation=0.8, n_periods=4, server_type=2, server_mu=1.5, server_sd=0.0,
                          dest_server=(2, 0.5, 0.0, 0.0)),
    # half of the OD pairs already have (several) routes: numbering continues, equal paths are not added twice
    "partial_routes": dict(nx=6, ny=6, seed=32, od_every=2, n_od=24, trips_per_hour=500, horizon=2700.0, keep_routes=0.5,
                           alt_routes=2, hist_variation=0.5, n_periods=3, server_type=2, server_mu=1.2, server_sd=0.0,
                           dest_server=(2, 0.4, 0.0, 0.0)),
    # FIFO link-time table (rows non-decreasing over the periods) that covers every label: the frontier kernel certifies
    # its trees itself instead of falling back to the exact kernel
    "fifo_table": dict(nx=7, ny=7, seed=33, od_every=2, n_od=36, trips_per_hour=300, horizon=3600.0, keep_routes=0.0,
                       hist_variation=0.7, hist_fifo=True, n_periods=4, server_type=2, server_mu=1.5, server_sd=0.0,
                       dest_server=(2, 0.5, 0.0, 0.0), len_lo=100, len_hi=300),
}


def parse_routes_file(path):
    txt = open(path).read()
    out = []
    for m in re.finditer(r"\{\s*(\d+)\s+(\d+)\s+(\d+)\s+(\d+)\s*\{([^}]*)\}\s*\}", txt):
        rid, o, d, n = (int(m.group(i)) for i in range(1, 5))
        links = [int(x) for x in m.group(5).split()]
        assert len(links) == n
        out.append((rid, o, d, links))
    return out


def oracle_route_fn(spec):
    n_nodes, n_links, up, dn, pro, T, P, pl = mzroutes.graph_arrays(spec)
    og = orc.OracleGraph(orc.GraphSpec(n_nodes, n_links, up, dn, None, pro))
    og.set_linktimes(P, pl, T)

    def fn(roots, entries, dest_off, dests):
        off, links = [0], []
        for t, (r, e) in enumerate(zip(roots, entries)):
            lp, npd, _, _, _ = og.tree(int(r), float(e))
            for d in dests[dest_off[t]:dest_off[t + 1]]:
                links.extend(og.route(int(r), lp, npd, int(d)).tolist())
                off.append(len(links))
        return np.array(off, np.int64), np.array(links, np.int32)

    return fn


def reference_routes(spec, tmp_path):
    r = refrun.run_reference(spec, workdir=str(tmp_path), calc_paths=1, keep=True)
    return r, parse_routes_file(os.path.join(str(tmp_path), "routes.dat"))


@pytest.mark.skipif(not refrun.have_ref(), reason="oracle/_ref/mezzo_ref_run not built (reference tree absent)")
@pytest.mark.parametrize("name", sorted(CASES))
def test_route_sets_match_the_reference(built, tmp_path, name):
    spec = synth.sim_grid(**CASES[name])
    r, ref_routes = reference_routes(spec, tmp_path)
    got = mzroutes.shortest_paths_all(spec, oracle_route_fn(spec))
    assert len(got) > len(spec["routes"])
    assert got == ref_routes  # ids, OD pairs, link sequences and multimap order
    # the simulation on the generated route sets (route choice among them) reproduces the reference run
    spec2 = dict(spec, routes=got)
    f = flatten.FlatNet(spec2)
    n = len(r["trips"]["time"])
    assert np.array_equal(r["trips"]["time"], f.trip_time[:n])
    assert np.array_equal(r["trips"]["route"], f.route_ids[f.trip_route[:n]])
    o = orc.oracle_sim(f)
    a, b = np.sort(r["exits"], order=["vid", "entry"]), np.sort(o["exits"], order=["vid", "entry"])
    assert len(a) == len(b) and np.array_equal(a["vid"], b["vid"]) and np.array_equal(a["lid"], f.link_ids[b["link"]])
    assert np.array_equal(a["entry"], b["entry"]) and np.array_equal(a["exit"], b["exit"])


@pytest.mark.gpu
@pytest.mark.parametrize("name", sorted(CASES))
def test_route_sets_gpu_equal_oracle_driver(built, name):
    """The same driver on the engine (mz_sssp_routes, both kernels) gives the identical route list."""
    import mezzo_b200
    spec = synth.sim_grid(**CASES[name])
    want = mzroutes.shortest_paths_all(spec, oracle_route_fn(spec))
    for mode in (mezzo_b200.MZ_SSSP_EXACT, mezzo_b200.MZ_SSSP_AUTO):
        fn = mzroutes.gpu_route_fn(spec, mode=mode)
        assert mzroutes.shortest_paths_all(spec, fn) == want
        st = fn.graph.last_stats()
        if name == "fifo_table" and mode == mezzo_b200.MZ_SSSP_AUTO:
            assert st["fast_steps"] > 0 and st["n_exact_redo"] < 0.5 * (len(want) + 1)  # mostly certified on the device
        fn.graph.close()
