"""CPU tests (no GPU): the C restatement of the link-update event loop (oracle/sim_oracle.c) and the host trip
pre-pass (mezzo_b200/flatten.py) against
  (1) committed golden vectors produced by the unmodified reference (tests/golden/sim_golden.npz), and
  (2) the compiled reference itself (oracle/_ref/mezzo_ref_run) when present.
Bar: BIT-EXACT per-(vehicle, link) entry/exit times and OD rows — the CPU port shares glibc's libm with the reference."""
import os

import numpy as np
import pytest

import orc
import refrun
import simcases
from mezzo_b200 import flatten


@pytest.fixture(scope="module")
def golden():
    return np.load(os.path.join(os.path.dirname(__file__), "golden", "sim_golden.npz"))


def sorted_exits(ex):
    return np.sort(ex, order=["vid", "entry"])


@pytest.mark.parametrize("name", simcases.GOLDEN)
def test_oracle_matches_golden(built, golden, name):
    f = flatten.FlatNet(simcases.make(name))
    # host pre-pass: trip start times and routes are what ODaction::execute produced in the reference
    gt = golden[f"{name}_trip_time"]
    assert np.array_equal(gt, f.trip_time[:len(gt)])
    assert np.array_equal(golden[f"{name}_trip_route"], f.route_ids[f.trip_route[:len(gt)]])
    o = orc.oracle_sim(f)
    ex = sorted_exits(o["exits"])
    assert len(ex) == len(golden[f"{name}_vid"])
    assert np.array_equal(ex["vid"], golden[f"{name}_vid"])
    assert np.array_equal(f.link_ids[ex["link"]], golden[f"{name}_lid"])
    assert np.array_equal(ex["entry"], golden[f"{name}_entry"])
    assert np.array_equal(ex["exit"], golden[f"{name}_exit"])
    assert o["stats"]["end_time"] == float(golden[f"{name}_currenttime"])
    # OD rows: o, d, vid, start, end, end-start, meters, route, switched
    rows = golden[f"{name}_arrivals"]
    arr = o["arrivals"]
    assert (arr >= 0).sum() == len(rows)
    vid = rows[:, 2].astype(int) - 1
    assert np.array_equal(arr[vid], rows[:, 4])
    assert np.array_equal(f.trip_time[vid], rows[:, 3])


@pytest.mark.parametrize("name", sorted(simcases.CASES))
def test_superstep_schedule_emulation_matches_oracle(built, name):
    """The product's per-event code and window/tie logic (mezzo_b200/csrc/sim_core.h + sim_host.h) compiled for the
    host (tests/emu) must reproduce the sequential event order bit for bit — including the tie-heavy networks."""
    f = flatten.FlatNet(simcases.make(name))
    o = orc.oracle_sim(f)
    e = orc.emu_sim(f)
    oe = sorted_exits(o["exits"])
    assert len(oe) == len(e["exits"])
    for k in ("vid", "link", "entry", "exit"):
        assert np.array_equal(oe[k], e["exits"][k]), k
    assert np.array_equal(o["arrivals"], e["arrivals"])
    assert np.array_equal(o["linktimes"], e["linktimes"])
    for k in ("n_traversals", "n_exits", "n_arrived", "end_time") + (() if name in simcases.SIGNALS else ("n_events",)):
        assert o["stats"][k] == e["stats"][k], k


def test_threaded_engine_build_matches_oracle(built):
    """The engine build (sim_host.h: sim_build) cuts its passes over trips, turnings, actions and nodes into pieces for
    worker threads only above size thresholds no golden case reaches. A quarter of the target network (255 k links, 67.5 k
    nodes, 767 k turnings, 292 k trips in the first 420 s) crosses all of them: the host emulation built that way must
    still reproduce the sequential oracle bit for bit."""
    from mezzo_b200 import synth
    horizon = 420.0
    spec = synth.sim_grid(nx=250, ny=250, seed=42, od_every=5, n_od=50_000, total_trips=int(round(2_500_000 * horizon / 3600.0)),
                          horizon=horizon, connector_len=250, od_radius=4, n_periods=1)
    f = flatten.FlatNet(spec)
    assert f.struct.n_trips > (1 << 18) and f.struct.n_turnings > (1 << 18) and f.struct.n_nodes > (1 << 16)
    o = orc.oracle_sim(f)
    e = orc.emu_sim(f)
    oe = sorted_exits(o["exits"])
    assert len(oe) == len(e["exits"]) > 500_000
    for k in ("vid", "link", "entry", "exit"):
        assert np.array_equal(oe[k], e["exits"][k]), k
    assert np.array_equal(o["arrivals"], e["arrivals"])
    assert np.array_equal(o["linktimes"], e["linktimes"])
    for k in ("n_traversals", "n_exits", "n_arrived", "end_time", "n_events"):
        assert o["stats"][k] == e["stats"][k], k


@pytest.mark.skipif(not refrun.have_ref(), reason="oracle/_ref/mezzo_ref_run not built (reference tree absent)")
@pytest.mark.parametrize("name", ["two_lane_lookback3", "const_servers_dynamit", "det_ties_grid", "signals_det", "signals_stoch"])
def test_oracle_matches_compiled_reference(built, name):
    spec = simcases.make(name)
    r = refrun.run_reference(spec)
    f = flatten.FlatNet(spec)
    o = orc.oracle_sim(f)
    a, b = sorted_exits(r["exits"]), sorted_exits(o["exits"])
    assert len(a) == len(b)
    assert np.array_equal(a["vid"], b["vid"]) and np.array_equal(a["lid"], f.link_ids[b["link"]])
    assert np.array_equal(a["entry"], b["entry"]) and np.array_equal(a["exit"], b["exit"])
    assert float(r["info"]["currenttime"]) == o["stats"]["end_time"]
    assert int(r["info"]["n_arrivals"]) == o["stats"]["n_arrived"]


@pytest.mark.skipif(not refrun.have_ref(), reason="oracle/_ref/mezzo_ref_run not built (reference tree absent)")
def test_reference_reproduces_its_own_golden_outputs(tmp_path):
    """Config 1: the bundled SF network — the harness must reproduce the reference's 13 ExpectedOutputs byte for byte."""
    import shutil
    import subprocess
    src = "/root/reference/branches/busmezzo/tests/networks/SFnetwork"
    if not os.path.isdir(src):
        pytest.skip("reference fixture not present")
    d = tmp_path / "sf"
    shutil.copytree(src, d)
    os.makedirs(d / "output", exist_ok=True)
    os.makedirs(d / "mz", exist_ok=True)
    subprocess.run([refrun.REF_RUN, "masterfile.mezzo", "42", "mz/", "save"], cwd=d, stdout=subprocess.DEVNULL,
                   stderr=subprocess.DEVNULL, check=True)
    exp = sorted(x for x in os.listdir(d / "ExpectedOutputs") if x.endswith(".dat"))
    assert len(exp) == 13
    for name in exp:
        got = d / name if (d / name).exists() else d / "output" / name
        assert open(got, "rb").read() == open(d / "ExpectedOutputs" / name, "rb").read(), name
    info = dict(l.strip().split("=") for l in open(d / "mz" / "info.txt"))
    assert float(info["currenttime"]) == 5400.1  # BT/integrationtest/test_integration.cpp:80-95
