"""Generates tests/golden/bus_sf_golden.json: the reference's own BusMezzo fixture (tests/networks/SFnetwork: 15 links, 4 bus
lines, 4 stops, 132 scheduled trips, passengers, no car demand) run by the UNMODIFIED reference (oracle/_ref/mezzo_ref_run;
the run reproduces the fixture's 13 ExpectedOutputs byte for byte), with what the transit hooks need
taken by compile-time macros (oracle/ref_hook_link.h, ref_hook_busline.h):
  * the network, flattened by this repo's readers (mezzo_b200/files.py) — a re-encoding of the fixture's input files;
  * every dispatched bus trip: time it enters the network, line, route, stops (link, position), vehicle length;
  * every stop visit with FULL precision: time the bus reaches the stop (Link::enter_veh / Busstop::execute booked it) and
    time it leaves (Busstop::calc_exiting_time — the transit model's answer, which the tests replay as the host model);
  * every link exit (link, entry time, exit time).
A second fixture, bus_sf_mixed_golden.json, is the same scenario with car demand switched on (main(MIXED_RATES, ...)).
Authoring container only:   make -C oracle ref && python tests/golden/make_bus_golden.py
"""
import json
import os
import shutil
import subprocess
import sys
import tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
import refrun  # noqa: E402
from mezzo_b200 import files  # noqa: E402

SRC = "/root/reference/branches/busmezzo/tests/networks/SFnetwork"
STOP_DT = np.dtype([("vid", "<i4"), ("trip", "<i4"), ("stop", "<i4"), ("link", "<i4"), ("enter", "<f8"), ("exit", "<f8")])
DISPATCH_DT = np.dtype([("vid", "<i4"), ("trip", "<i4"), ("time", "<f8"), ("length", "<f8")])
EXIT_DT = np.dtype([("vid", "<i4"), ("link", "<i4"), ("entry", "<f8"), ("exit", "<f8")])


MIXED_RATES = (240, 120, 160, 120)  # veh/h for the fixture's four OD pairs (all 0 in the fixture itself)
HEAVY_RATES = (500, 300, 400, 300)  # gridlock: 1 970 vehicles on links at the end, shock-wave exit-time rewrites that include buses


def main(car_rates=None, out_name="bus_sf_golden.json"):
    """car_rates=None: the fixture as it is (buses only). Otherwise the same scenario with CAR DEMAND on its four OD pairs —
    the one change made to a copy of the fixture's demand.dat — so that cars of its five-class vehicle mix (6 m cars, 16 m
    trucks) and buses share links, queues and origins: tests/golden/bus_sf_mixed_golden.json, bus_sf_heavy_golden.json."""
    wd = os.path.join(tempfile.mkdtemp(prefix="mzbus_"), "sf")
    shutil.copytree(SRC, wd)
    os.makedirs(os.path.join(wd, "output"), exist_ok=True)
    os.makedirs(os.path.join(wd, "mz"), exist_ok=True)
    if car_rates:
        lines = open(os.path.join(wd, "demand.dat")).read().split("\n")
        k = 0
        for i, ln in enumerate(lines):
            t = ln.replace("{", " ").replace("}", " ").split()
            if ln.strip().startswith("{") and len(t) == 3:
                lines[i] = "{\t%s\t%s\t%g\t}" % (t[0], t[1], car_rates[k])
                k += 1
        assert k == len(car_rates)
        open(os.path.join(wd, "demand.dat"), "w").write("\n".join(lines))
    subprocess.run([refrun.REF_RUN, "masterfile.mezzo", "42", "mz/", "save"], cwd=wd, stdout=subprocess.DEVNULL,
                   stderr=subprocess.DEVNULL, check=True)
    for name in sorted(os.listdir(os.path.join(wd, "ExpectedOutputs"))):  # the run must be the reference's golden run
        if name.endswith(".dat") and not car_rates:
            got = os.path.join(wd, name) if os.path.exists(os.path.join(wd, name)) else os.path.join(wd, "output", name)
            assert open(got, "rb").read() == open(os.path.join(wd, "ExpectedOutputs", name), "rb").read(), name
    m = files.read_master(os.path.join(wd, "masterfile.mezzo"))
    p = lambda k: os.path.join(wd, m[k])  # noqa: E731
    servers, nodes, sdfuncs, links = files.read_network(p("network"))
    P, pl, hist = files.read_linktimes(p("histtimes"))
    ods, slices = files.read_demand(p("demand"))
    spec = dict(servers=servers, nodes=nodes, sdfuncs=sdfuncs, links=links, turnings=files.read_turnings(p("turnings")), ods=ods,
                slices=slices, routes=files.read_routes(p("routes")), incident_sdfuncs=[], incidents=[],
                # (without car demand the car mix of vehiclemix.dat is never drawn from)
                vtypes=files.read_vtypes(p("vehicletypes")) if car_rates else [(4, "Buses", 1.0, 15.0)],
                n_periods=P, periodlength=pl, histtimes={str(k): v for k, v in hist.items()}, params=files.read_parameters(p("parameters")),
                serverrates=[], signals=[], starttime=m["starttime"], stoptime=m["stoptime"], randseed=42)
    trips = {}
    for line in open(os.path.join(wd, "mz", "bustrips.txt")):
        t = line.split()
        tid = int(t[t.index("trip") + 1])
        trips[tid] = dict(trip=tid, line=int(t[t.index("line") + 1]), length=float(t[t.index("length") + 1]),
                          starttime=float(t[t.index("starttime") + 1]), origin=int(t[t.index("origin") + 1]),
                          dest=int(t[t.index("dest") + 1]), route=int(t[t.index("route") + 1]),
                          links=[int(x) for x in t[t.index("links") + 1:t.index("stops")]],
                          stops=[[int(x.split(":")[0]), int(x.split(":")[1]), float(x.split(":")[2])] for x in t[t.index("stops") + 1:]])
    dispatch = np.fromfile(os.path.join(wd, "mz", "busdispatch.bin"), DISPATCH_DT)
    stops = np.fromfile(os.path.join(wd, "mz", "stops.bin"), STOP_DT)
    exits = np.fromfile(os.path.join(wd, "mz", "exits.bin"), EXIT_DT)
    info = dict(l.strip().split("=") for l in open(os.path.join(wd, "mz", "info.txt")))
    out = dict(spec=spec,
               # in dispatch (event) order; the vehicle length is the one the dispatch hook saw (bus vehicles are shared
               # between trips and reset when a trip ends — what the trip's vehicle holds after the run is not it)
               dispatched=[dict(trips[int(d["trip"])], time=float(d["time"]), length=float(d["length"])) for d in dispatch],
               visits=[[int(s["trip"]), int(s["stop"]), int(s["link"]), float(s["enter"]), float(s["exit"])] for s in stops],
               exits=[[int(e["link"]), float(e["entry"]), float(e["exit"])] for e in exits],
               n_transit_init_steps=int(info["n_transit_init_steps"]),
               currenttime=float(info["currenttime"]), n_on_links=int(info["n_on_links"]))
    if car_rates:  # every car the reference generated: vehicle id, time, route, the length of the class it drew
        tr = np.fromfile(os.path.join(wd, "mz", "trips.bin"), refrun.TRIP_DT)
        out["car_trips"] = [[int(t["vid"]), float(t["time"]), int(t["route"]), float(t["length"])] for t in tr]
        out["exit_vids"] = [int(e["vid"]) for e in exits]  # (-1 = a bus)
    if out_name is None:  # (tests/fuzz_sf.py: the run's record, not a fixture)
        return out
    with open(os.path.join(HERE, out_name), "w") as f:
        json.dump(out, f)
    print(out_name, len(out["dispatched"]), "bus trips,", len(out["visits"]), "stop visits,", len(out["exits"]), "link exits",
          len(out.get("car_trips", [])), "car trips")


if __name__ == "__main__":
    main()
    main(MIXED_RATES, "bus_sf_mixed_golden.json")
    main(HEAVY_RATES, "bus_sf_heavy_golden.json")
