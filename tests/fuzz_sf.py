"""TEST INFRASTRUCTURE (CPU, not collected by pytest; needs /root/reference): the reference's BusMezzo fixture (SFnetwork) with
RANDOM car demand on its four OD pairs — free flow to gridlock — run by the unmodified reference (tests/golden/
make_bus_golden.py), then by the C oracle and by the engine's own code compiled for the host with the reference's stop exit
times replayed as the transit model: every link exit (cars, trucks, buses) and every stop arrival must equal the reference's.
usage: python tests/fuzz_sf.py [seed] [count]"""
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
sys.path.insert(0, HERE)
sys.path.insert(0, os.path.join(HERE, "golden"))
import numpy as np

import make_bus_golden
import orc
import test_stops
from mezzo_b200.sim import run_with_stops

rng = np.random.default_rng(int(sys.argv[1]) if len(sys.argv) > 1 else 0)
N = int(sys.argv[2]) if len(sys.argv) > 2 else 10
bad = 0
for it in range(N):
    rates = [int(r) for r in rng.choice([0, 30, 60, 120, 240, 400, 600, 900], 4)]
    if not any(rates):
        rates[0] = 60
    g = make_bus_golden.main(tuple(rates), None)
    f, cars, trip_of_veh, ref_vis, ref_exits, _ = test_stops.sf_mixed_fixture(g)
    ref = np.array(g["car_trips"]).reshape(-1, 4)
    n = len(ref)
    ok_trips = np.array_equal(cars["time"][:n], ref[:, 1]) and np.array_equal(cars["length"][:n], ref[:, 3])
    sim = orc.EmuSimulation(f)
    visits = run_with_stops(sim, lambda v: ref_vis[(trip_of_veh[int(v["veh"])], int(v["stop"]))][1], 1.0)
    ex, st = sim.exit_log(), sim.stats()
    sim.close()
    got = np.array(sorted((int(e["link"]), float(e["entry"]), float(e["exit"])) for e in ex), np.float64).reshape(-1, 3)
    mine = {(trip_of_veh[int(v["veh"])], int(v["stop"])): float(v["t_arrive"]) for v in visits}
    ok_emu = got.shape == ref_exits.shape and np.array_equal(got, ref_exits) and all(mine.get(k) == e for k, (e, _l) in ref_vis.items()) \
        and st["end_time"] == g["currenttime"]
    veh_of_trip = {t: v for v, t in trip_of_veh.items()}
    dwell = np.zeros(len(f.stop_link))
    for (tr, k), (enter, leave) in ref_vis.items():
        dwell[f.trip_stop_off[veh_of_trip[tr]] + k] = leave - enter
    o = orc.oracle_sim(f, dwell)
    og = np.array(sorted((int(e["link"]), float(e["entry"]), float(e["exit"])) for e in o["exits"]), np.float64).reshape(-1, 3)
    ok_or = og.shape == ref_exits.shape and np.array_equal(og[:, 0], ref_exits[:, 0]) and np.allclose(og[:, 1:], ref_exits[:, 1:], rtol=1e-12, atol=0)
    # (vehicles on links: the reference's figure is the sum of its queue sizes, which leaves out the buses standing at a stop
    # — or lost at one: Busstop::execute gives a bus up when the queue it should re-enter is full, B/busline.cpp:1240-1248 —
    # so the engine's count is compared with the oracle's)
    ok_emu = ok_emu and all(st[k] == o["stats"][k] for k in ("n_traversals", "n_exits", "n_arrived"))
    good = ok_trips and ok_emu and ok_or
    bad += not good
    print(it, "OK " if good else "MISMATCH", "rates", rates, "cars", n, "exits", len(ref_exits), "visits", len(ref_vis), "on links", g["n_on_links"],
          "" if good else f"trips {ok_trips} emulation {ok_emu} oracle {ok_or}", flush=True)
print("bad", bad)
