"""TEST INFRASTRUCTURE (CPU, not collected by pytest): randomised parity sweep of the link update.
Random small networks with random feature mixes (signals, server-rate changes, incidents, demand slices, vehicle classes, two-lane links, look-back, congestion levels,
route choice, server types) — the engine's own per-event code compiled for the host (tests/emu) against the CPU oracle, and
every fourth network also against the compiled reference.   usage: python tests/fuzz_parity.py [seed] [count]"""
import sys, time
import os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__)))); sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import numpy as np
import orc, refrun
from mezzo_b200 import flatten, synth
rng = np.random.default_rng(int(sys.argv[1]) if len(sys.argv) > 1 else 0)
N = int(sys.argv[2]) if len(sys.argv) > 2 else 20
bad = 0
for it in range(N):
    det = bool(rng.integers(2))
    kw = dict(nx=int(rng.integers(4, 8)), ny=int(rng.integers(4, 8)), seed=int(rng.integers(1000)), od_every=2,
              n_od=int(rng.integers(10, 40)), trips_per_hour=int(rng.integers(200, 2500)), horizon=float(rng.choice([600.0, 1200.0, 1800.0])),
              two_lane_prob=float(rng.choice([0.0, 0.3, 0.6])), lookback=int(rng.choice([2, 4, 20])),
              len_lo=int(rng.choice([40, 60, 150, 300])), connector_len=int(rng.choice([30, 50, 150, 250])),
              signals=int(rng.choice([0, 0, 4, 10])), rate_changes=int(rng.choice([0, 0, 20])),
              alt_routes=int(rng.choice([1, 1, 3])), hist_variation=float(rng.choice([0.0, 0.5])), n_periods=int(rng.choice([2, 4])))
    kw["len_hi"] = kw["len_lo"] + int(rng.choice([0, 50, 300]))
    kw.update(incidents=int(rng.choice([0, 0, 2])), demand_slices=int(rng.choice([0, 0, 2])))
    if rng.integers(3) == 0:  # several vehicle classes (the per-trip draw of Vtypes::random_vtype)
        kw["vehicle_classes"] = [(float(p), float(ln)) for p, ln in zip(rng.uniform(0.05, 0.5, 3), rng.choice([4.0, 6.0, 9.5, 12.0, 18.0], 3, replace=False))]
    if det:
        kw.update(server_type=2, server_mu=float(rng.choice([1.0, 1.5, 2.0])), server_sd=0.0, dest_server=(2, float(rng.choice([0.3, 0.5, 1.0])), 0.0, 0.0))
    else:
        pst = int(rng.choice([1, 1, 3, 4]))
        kw.update(param_server_type=pst)
        if pst == 3: kw.update(server_delay=0.5, server_mu=1.0, dest_server=(1, 0.3, 0.1, 0.2))
        if pst == 4: kw.update(server_sd=6.0, dest_server=(1, 0.5, 6.0, 0.0))
    spec = synth.sim_grid(**kw)
    f = flatten.FlatNet(spec)
    o = orc.oracle_sim(f); e = orc.emu_sim(f)
    b, c = np.sort(o["exits"], order=["vid", "entry"]), e["exits"]
    same = len(c) == len(b) and np.array_equal(c["vid"], b["vid"]) and np.array_equal(c["exit"], b["exit"]) and np.array_equal(o["arrivals"], e["arrivals"]) and np.array_equal(o["linktimes"], e["linktimes"])
    refok = ""
    if it % 4 == 0:
        r = refrun.run_reference(spec)
        a = np.sort(r["exits"], order=["vid", "entry"])
        refok = " ref==oracle %s" % (len(a) == len(b) and np.array_equal(a["exit"], b["exit"]))
    if not same or "False" in refok: bad += 1
    print(it, "OK " if same else "MISMATCH", len(b), "exits", {k: kw[k] for k in ("signals", "rate_changes", "incidents", "demand_slices", "two_lane_prob", "lookback", "trips_per_hour")}, "mix" if "vehicle_classes" in kw else "", "det" if det else "stoch", refok, flush=True)
print("bad", bad)
