"""Transit overlay hooks: type-4 vehicles leave the link queue for a stop visit (B/link.cpp:489-516)
and come back through Q::enter_veh when the host's transit model lets them go (Busstop::execute, B/busline.cpp:1196-1263).
The engine side is mz_sim_run_until / mz_sim_stop_arrivals / mz_sim_stop_departures; the transit model stays on the host —
here a dwell-time table drives it. The oracle (oracle/sim_oracle.c) runs the same network sequentially with the same table.
Parity of these two reference pieces is UNPINNED (no BusMezzo run of the compiled reference backs the oracle's restatement).
CPU: oracle vs the host emulation of the engine code driven through the same co-simulation loop. GPU: the C ABI."""
import numpy as np
import pytest

import orc
import simcases
from mezzo_b200 import flatten
from mezzo_b200.sim import run_with_stops

MIN_DWELL = 15.0


def with_stops(name, every=5, seed=0, same_link_prob=0.3):
    """A simcases network in which every `every`-th trip is a bus with 1-4 stops on links of its route (never the first
    link: the bus enters that one from its origin), some of them two on one link; dwell[n_stops] in [15, 60) s."""
    f = flatten.FlatNet(simcases.make(name))
    rng = np.random.default_rng(seed)
    off, links, pos = [0], [], []
    for v in range(len(f.trip_time)):
        r = f.trip_route[v]
        rl = f.route_links[f.route_off[r] + 1:f.route_off[r + 1]]
        if v % every == 0 and len(rl) >= 2:
            k = int(rng.integers(1, min(4, len(rl)) + 1))
            for i in sorted(rng.choice(len(rl), size=k, replace=False)):
                l = int(rl[i])
                length = float(f.link_length[l])
                p = rng.uniform(0.05, 0.95, size=2 if rng.random() < same_link_prob else 1) * length
                for x in sorted(p):
                    links.append(l)
                    pos.append(float(x))
        off.append(len(links))
    f.set_stops(off, links, pos)
    dwell = rng.uniform(MIN_DWELL, 60.0, size=len(links))
    return f, dwell


def drive(sim, f, dwell, slice_s=MIN_DWELL):
    def depart(v):
        return v["t_arrive"] + dwell[f.trip_stop_off[v["veh"]] + v["stop"]]
    visits = run_with_stops(sim, depart, slice_s)
    out = dict(exits=np.sort(sim.exit_log(), order=["vid", "entry"]), arrivals=sim.arrivals(), linktimes=sim.linktimes(),
               stats=sim.stats(), visits=np.sort(visits, order=["veh", "stop"]))
    sim.close()
    return out


def compare(g, o, exact):
    ov = np.sort(o["visits"], order=["veh", "stop"])
    oe = np.sort(o["exits"], order=["vid", "entry"])
    assert len(ov) > 20, "the case must exercise stops"
    assert len(g["visits"]) == len(ov)
    for k in ("veh", "stop", "link"):
        assert np.array_equal(g["visits"][k], ov[k])
    assert len(g["exits"]) == len(oe)
    assert np.array_equal(g["exits"]["vid"], oe["vid"]) and np.array_equal(g["exits"]["link"], oe["link"])
    if exact:
        for k in ("t_enter", "t_arrive"):
            assert np.array_equal(g["visits"][k], ov[k])
        assert np.array_equal(g["exits"]["entry"], oe["entry"]) and np.array_equal(g["exits"]["exit"], oe["exit"])
        assert np.array_equal(g["arrivals"], o["arrivals"])
        assert np.array_equal(g["linktimes"], o["linktimes"], equal_nan=True)
    else:
        for k in ("t_enter", "t_arrive"):
            np.testing.assert_allclose(g["visits"][k], ov[k], rtol=1e-6, atol=0)
        np.testing.assert_allclose(g["exits"]["exit"], oe["exit"], rtol=1e-6, atol=0)
    for k in ("n_traversals", "n_exits", "n_arrived", "n_stop_visits", "n_bus_refused"):
        assert g["stats"][k] == o["stats"][k], k


CASES = ["det_two_lane_congested", "det_servers", "det_gridlock", "const_servers_linear"]


@pytest.mark.parametrize("name", CASES)
def test_stops_emulation_vs_oracle(name):
    f, dwell = with_stops(name)
    o = orc.oracle_sim(f, dwell)
    assert o["stats"]["n_stop_visits"] == len(f.stop_link) or o["stats"]["n_arrived"] < len(f.trip_time)
    g = drive(orc.EmuSimulation(f), f, dwell)
    compare(g, o, exact=True)


def test_buses_change_the_traffic():
    """the hooks are not a no-op: with stops the same network produces different exit times than without"""
    f, dwell = with_stops("det_servers")
    o = orc.oracle_sim(f, dwell)
    plain = orc.oracle_sim(flatten.FlatNet(simcases.make("det_servers")))
    assert len(o["visits"]) > 0
    assert not np.array_equal(np.sort(o["exits"], order=["vid", "entry"])["exit"][:len(plain["exits"])],
                              np.sort(plain["exits"], order=["vid", "entry"])["exit"][:len(o["exits"])])


def test_slice_length_does_not_matter():
    f, dwell = with_stops("det_two_lane_congested", seed=3)
    a = drive(orc.EmuSimulation(f), f, dwell, slice_s=MIN_DWELL)
    b = drive(orc.EmuSimulation(f), f, dwell, slice_s=4.0)
    for k in ("vid", "link", "entry", "exit"):
        assert np.array_equal(a["exits"][k], b["exits"][k])
    assert np.array_equal(a["visits"], b["visits"])


def test_refusals():
    f, dwell = with_stops("det_servers")
    sim = orc.EmuSimulation(f)
    sim.run_until(600.0)
    vis = sim.stop_arrivals()
    assert len(vis) > 0
    with pytest.raises(RuntimeError, match="precedes the time the engine has run to"):
        sim.stop_departures([int(vis[0]["veh"])], [100.0])
    car = next(v for v in range(len(f.trip_time)) if f.trip_stop_off[v] == f.trip_stop_off[v + 1])
    with pytest.raises(RuntimeError, match="not at a stop"):
        sim.stop_departures([car], [700.0])
    with pytest.raises(RuntimeError, match="already run"):
        sim.run_until(100.0)
    sim.close()
    bad = flatten.FlatNet(simcases.make("det_servers"))
    off = np.zeros(len(bad.trip_time) + 1, np.int32)
    off[1:] = 1
    r = bad.trip_route[0]
    bad.set_stops(off, [int(bad.route_links[bad.route_off[r]])], [1.0])  # a stop on the first link of the route
    with pytest.raises(RuntimeError, match="route order"):
        orc.EmuSimulation(bad)


@pytest.mark.gpu
@pytest.mark.parametrize("kernel", [1, 2, 3])
@pytest.mark.parametrize("name", ["det_two_lane_congested", "det_servers"])
def test_stops_gpu_vs_oracle(name, kernel):
    import mezzo_b200
    f, dwell = with_stops(name)
    o = orc.oracle_sim(f, dwell)
    sim = mezzo_b200.Simulation(f)
    sim.set_option(mezzo_b200.MZ_SIM_OPT_KERNEL, kernel)
    compare(drive(sim, f, dwell), o, exact=True)


@pytest.mark.gpu
def test_stops_gpu_stochastic_servers():
    import mezzo_b200
    f, dwell = with_stops("congested_short", every=4)
    o = orc.oracle_sim(f, dwell)
    compare(drive(mezzo_b200.Simulation(f), f, dwell), o, exact=False)
