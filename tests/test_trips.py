"""The native host pre-pass for the trip table (mz_trips_generate, mezzo_b200/csrc/trips.cu) against
its Python restatement of the same reference code (flatten.generate_trips_py). The reference itself pins both:
tests/test_oracle_sim.py and tests/test_routes.py compare FlatNet's trip table — now produced natively — with the
ODaction log of the compiled reference (times, chosen routes)."""
import numpy as np
import pytest

import simcases
from mezzo_b200 import flatten


@pytest.mark.parametrize("name", sorted(simcases.CASES))
def test_native_trip_table_equals_python_restatement(built, name):
    spec = simcases.make(name)
    f = flatten.FlatNet(spec)
    nt, ods, ori, oidx = f._trip_args
    ref = flatten.generate_trips_py(
        spec, nt, ods, ori, oidx, list(spec["vtypes"]),
        route_hist=[f.link_hist[f.route_links[f.route_off[r]:f.route_off[r + 1]]] for r in range(len(f.route_ids))])
    got = (f.trip_time, f.trip_origin, f.trip_route, f.trip_length, f.trip_od_sorted, f.n_odpairs, f.trip_od,
           f.trip_prev_time, f.phantom_final)
    assert len(ref) == len(got)
    for a, b in zip(ref, got):
        assert np.array_equal(np.asarray(a), np.asarray(b))
    assert len(f.trip_time) > 1000


@pytest.mark.parametrize("name", ["vehicle_mix_det", "vehicle_mix_stoch"])
def test_vehicle_classes_match_the_reference(built, name):
    """Several vehicle classes: the class of every trip (seen through the vehicle length the reference's ODaction hands to
    Vehicle::init, logged by the unmodified reference into tests/golden/sim_golden.npz) — one network-wide Random stream,
    one draw per trip in global event order (B/od.cpp:79, B/vtypes.h:61-71), native draw == Python restatement == reference."""
    import os
    golden = np.load(os.path.join(os.path.dirname(__file__), "golden", "sim_golden.npz"))
    spec = simcases.make(name)
    f = flatten.FlatNet(spec)
    ref = golden[f"{name}_trip_length"]
    n = len(ref)  # (the reference stops at the first event beyond the horizon; the pre-pass books one more per OD pair)
    assert n > 5000 and n <= len(f.trip_length)
    assert np.array_equal(f.trip_length[:n], ref)
    assert len(np.unique(ref)) == len(spec["vtypes"]) > 1
    py, idx = flatten.trip_lengths(list(spec["vtypes"]), spec["randseed"], len(f.trip_length), native=False)
    assert np.array_equal(py, f.trip_length)
    share = np.bincount(idx) / len(idx)
    assert np.allclose(share, flatten.vtype_probabilities(spec["vtypes"]), atol=0.02)


def test_delete_bad_routes_prunes_like_the_reference(built, tmp_path):
    """delete_bad_routes=1: the route sets left by ODpair::delete_spurious_routes (routes.dat written back by the reference)
    and the route choice of the run that follows (ODaction log: time and route of every trip)."""
    import os
    import refrun
    import test_routes
    if not refrun.have_ref():
        pytest.skip("oracle/_ref/mezzo_ref_run not built")
    spec = simcases.make("route_pruning_det")
    r = refrun.run_reference(spec, workdir=str(tmp_path), keep=True)
    ref_routes = test_routes.parse_routes_file(os.path.join(str(tmp_path), "routes.dat"))
    f = flatten.FlatNet(spec)
    srt = [x for _, x in sorted(enumerate(spec["routes"]), key=lambda x: (x[1][1], x[1][2], x[0]))]
    kept = {spec["routes"][i][0] for i in np.nonzero(f.route_kept)[0]}
    assert 0 < len(kept) < len(spec["routes"])  # something was really pruned
    assert [x for x in srt if x[0] in kept] == ref_routes
    n = len(r["trips"]["time"])
    assert np.array_equal(r["trips"]["time"], f.trip_time[:n])
    assert np.array_equal(r["trips"]["route"], f.route_ids[f.trip_route[:n]])


def test_phantom_final_event_ends_the_run(built):
    """SURVEY H8 with demand slices: when the first event at/after the horizon is a MatrixAction (or the poll of an inactive
    ODaction), that is the one event Network::step still executes — no action of the network runs beyond the horizon. Checked
    on the host emulation of the engine against the compiled reference."""
    import orc
    import refrun
    spec = simcases.make("demand_slices_det")
    spec["slices"] = list(spec["slices"]) + [(spec["stoptime"], 1.0, [spec["slices"][0][2][0]])]  # a MatrixAction AT the horizon
    f = flatten.FlatNet(spec)
    assert f.phantom_final == (spec["stoptime"], -np.inf)
    e = orc.emu_sim(f)
    o = orc.oracle_sim(f)
    assert e["stats"]["end_time"] == spec["stoptime"] == o["stats"]["end_time"]
    if refrun.have_ref():
        r = refrun.run_reference(spec)
        assert float(r["info"]["currenttime"]) == spec["stoptime"]
        a, b = np.sort(r["exits"], order=["vid", "entry"]), np.sort(e["exits"], order=["vid", "entry"])
        assert len(a) == len(b) and np.array_equal(a["exit"], b["exit"])
