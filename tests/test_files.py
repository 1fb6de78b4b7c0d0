"""File formats either side of the hot paths (mezzo_b200/files.py): the readers against the input
files the tests hand to the UNMODIFIED reference, the writers byte for byte against the files the reference's
Network::writeall / writepathfile leave behind for the same run. CPU tests feed the writers from the host emulation of
the engine (tests/emu); the GPU test runs a scenario directory end to end (files.run_masterfile)."""
import filecmp
import os
import shutil

import numpy as np
import pytest

import orc
import refrun
import simcases
from mezzo_b200 import files, flatten, sim as mzsim, synth

OUT_FILES = ["output/linktimes.dat", "output/linktimes.dat.clean", "output/output.dat", "output/summary.dat", "routes.dat"]
# every printed digit must agree (the host emulation and the reference share glibc, so N(mu,sd) servers are exact here too;
# the GPU test below keeps to deterministic servers: CUDA libm may differ from glibc in the last ulp)
CASES = ["det_ties_small", "det_gridlock", "det_two_lane_congested", "route_choice_det", "congested_short", "poisson_demand",
         "rate_changes_det", "rate_changes_stoch", "route_pruning_det", "lognormal_servers", "loglogistic_servers", "signals_det", "signals_stoch",
         "demand_slices_det", "incidents_det", "vehicle_mix_det", "vehicle_mix_stoch"]


def test_readers_round_trip(tmp_path):
    spec = simcases.make("route_choice_det")
    refrun.write_inputs(spec, str(tmp_path), calc_paths=1)
    got, m = files.read_scenario(os.path.join(str(tmp_path), "masterfile.mezzo"), spec["randseed"])
    assert m["calc_paths"] == 1 and m["traveltime_alpha"] == 0.4 and m["linktimes"] == "output/linktimes.dat"
    for k in ("servers", "nodes", "sdfuncs", "links", "turnings", "routes", "n_periods", "periodlength", "starttime",
              "stoptime", "randseed"):
        a, b = got[k], spec[k]
        if isinstance(b, list):
            a, b = [tuple(x) for x in a], [tuple(x) for x in b]
        assert a == b, k
    assert [(o, d, float(r)) for o, d, r in got["ods"]] == [(o, d, float(r)) for o, d, r in spec["ods"]]
    assert got["histtimes"] == spec["histtimes"]
    assert [tuple(v) for v in got["vtypes"]] == [tuple(v) for v in spec["vtypes"]]
    for k, v in spec["params"].items():
        if k in got["params"]:
            assert got["params"][k] == v, k
    # and the flattened network is the same object the tests run
    fa, fb = flatten.FlatNet(got), flatten.FlatNet(spec)
    for name in ("link_length", "turn_in", "turn_out", "trip_time", "trip_route", "link_hist", "route_links"):
        assert np.array_equal(getattr(fa, name), getattr(fb, name)), name


def test_refuses_what_is_not_built(tmp_path):
    spec = simcases.make("freeflow_small")
    d = str(tmp_path)
    refrun.write_inputs(spec, d)
    open(os.path.join(d, "signal.dat"), "w").write("controls: 1\n{ 1 1 { 1 0.0 900.0 0.0 60.0 1 { 1 0.0 30.0 1 { 1 } } } }\n")
    sig_spec, _ = files.read_scenario(os.path.join(d, "masterfile.mezzo"), 42)
    assert sig_spec["signals"] == [(1, [(1, 0.0, 900.0, 0.0, 60.0, [(1, 0.0, 30.0, [1])])])]
    open(os.path.join(d, "signal.dat"), "w").write("controls: 0\n")
    # an incident WITH information broadcast (info_start / info_stop >= 0: en-route switching) is read but not flattened
    lid = spec["links"][0][0]
    open(os.path.join(d, "noincident.dat"), "w").write(
        "sdfuncs: 1\n{ 77 0 5.0 2.0 150.0 10.0 }\nincidents: 1\n{ %d 77 0.0 100.0 400.0 150.0 500.0 0 }\n" % lid)
    got, _ = files.read_scenario(os.path.join(d, "masterfile.mezzo"), 42)
    assert got["incidents"] == [(lid, 77, 0.0, 100.0, 400.0, 150.0, 500.0, 0)] and got["incident_sdfuncs"][0][0] == 77
    with pytest.raises(NotImplementedError):
        flatten.FlatNet(got)
    open(os.path.join(d, "noincident.dat"), "w").write("sdfuncs: 0\nincidents: 0\n")
    open(os.path.join(d, "virtuallinks.dat"), "w").write("virtuallinks: 1\n")
    with pytest.raises(NotImplementedError):
        files.read_scenario(os.path.join(d, "masterfile.mezzo"), 42)


def test_demand_slices_round_trip(tmp_path):
    """demand.dat "slices:" (Network::readrates, B/network.cpp:5544-5606) → spec["slices"] → the same trip table."""
    spec = simcases.make("demand_slices_det")
    assert len(spec["slices"]) == 4
    refrun.write_inputs(spec, str(tmp_path))
    got, _ = files.read_scenario(os.path.join(str(tmp_path), "masterfile.mezzo"), spec["randseed"])
    assert [(lt, sc, [tuple(r) for r in rates]) for lt, sc, rates in got["slices"]] == \
        [(lt, sc, [tuple(r) for r in rates]) for lt, sc, rates in spec["slices"]]
    fa, fb = flatten.FlatNet(got), flatten.FlatNet(spec)
    for name in ("trip_time", "trip_route", "trip_od", "trip_prev_time"):
        assert np.array_equal(getattr(fa, name), getattr(fb, name)), name
    assert fa.phantom_final == fb.phantom_final and fa.phantom_final[0] >= spec["stoptime"]


def test_serverrates_round_trip(tmp_path):
    spec = simcases.make("rate_changes_det")
    assert len(spec["serverrates"]) > 60
    refrun.write_inputs(spec, str(tmp_path))
    got, _ = files.read_scenario(os.path.join(str(tmp_path), "masterfile.mezzo"), spec["randseed"])
    assert [tuple(r) for r in got["serverrates"]] == [tuple(r) for r in spec["serverrates"]]
    fa, fb = flatten.FlatNet(got), flatten.FlatNet(spec)
    for name in ("chg_server", "chg_time", "chg_mu", "chg_sd"):
        assert np.array_equal(getattr(fa, name), getattr(fb, name)), name


def _reference_dir(spec, tmp_path, calc_paths=0):
    ref = os.path.join(str(tmp_path), "ref")
    refrun.run_reference(spec, workdir=ref, calc_paths=calc_paths, keep=True, save=True)
    return ref


def _write_all(spec, m, d, arrivals, lt_final):
    f = flatten.FlatNet(spec)
    files.write_routes(os.path.join(d, m["routes"]), spec["routes"])
    files.write_linktimes(os.path.join(d, m["linktimes"]), f.link_ids, lt_final, f.link_hist, m["traveltime_alpha"],
                          spec["periodlength"])
    files.write_linktimes(os.path.join(d, m["linktimes"]) + ".clean", f.link_ids, lt_final, f.link_hist, 1.0,
                          spec["periodlength"])
    rows = mzsim.od_rows(f, arrivals)
    nr = {}
    for r in spec["routes"]:
        nr[(r[1], r[2])] = nr.get((r[1], r[2]), 0) + 1
    files.write_summary(os.path.join(d, m["summary"]), [(o, dd, nr[(o, dd)]) for o, dd, _ in f.od_sorted if (o, dd) in nr], rows)
    files.write_output(os.path.join(d, m["output"]), rows)


@pytest.mark.skipif(not refrun.have_ref(), reason="oracle/_ref/mezzo_ref_run not built (reference tree absent)")
@pytest.mark.parametrize("name", CASES)
def test_writers_match_the_reference_files(built, tmp_path, name):
    spec = simcases.make(name)
    ref = _reference_dir(spec, tmp_path)
    mine = os.path.join(str(tmp_path), "mine")
    refrun.write_inputs(spec, mine)
    got, m = files.read_scenario(os.path.join(mine, "masterfile.mezzo"), spec["randseed"])
    got["routes"] = [r for _, r in sorted(enumerate(got["routes"]), key=lambda x: (x[1][1], x[1][2], x[0]))]
    f0 = flatten.FlatNet(got)
    got["routes"] = [r for r, k in zip(got["routes"], f0.route_kept) if k]  # what add_od_routes left in the routemap
    e = orc.emu_sim(flatten.FlatNet(got))
    _write_all(got, m, mine, e["arrivals"], e["linktimes_final"])
    for fn in OUT_FILES:
        assert filecmp.cmp(os.path.join(ref, fn), os.path.join(mine, fn), shallow=False), fn


@pytest.mark.gpu
@pytest.mark.skipif(not refrun.have_ref(), reason="oracle/_ref/mezzo_ref_run not built")
@pytest.mark.parametrize("name,calc_paths", [("det_ties_small", 0), ("route_choice_det", 0), ("fifo_routes", 1)])
def test_scenario_directory_end_to_end_gpu(built, tmp_path, name, calc_paths):
    """files.run_masterfile on a scenario directory = what mezzo_s does with it: same routes.dat written back, same
    link times (smoothed and clean), OD output and summary, byte for byte."""
    if name == "fifo_routes":
        spec = synth.sim_grid(nx=7, ny=7, seed=33, od_every=2, n_od=36, trips_per_hour=300, horizon=3600.0, keep_routes=0.0,
                              hist_variation=0.7, hist_fifo=True, n_periods=4, server_type=2, server_mu=1.5, server_sd=0.0,
                              dest_server=(2, 0.5, 0.0, 0.0), len_lo=100, len_hi=300)
    else:
        spec = simcases.make(name)
    ref = _reference_dir(spec, tmp_path, calc_paths=calc_paths)
    mine = os.path.join(str(tmp_path), "mine")
    refrun.write_inputs(spec, mine, calc_paths=calc_paths)
    stats, _ = files.run_masterfile(os.path.join(mine, "masterfile.mezzo"), spec["randseed"])
    assert stats["n_traversals"] > 0
    for fn in OUT_FILES:
        assert filecmp.cmp(os.path.join(ref, fn), os.path.join(mine, fn), shallow=False), fn
    shutil.rmtree(ref, ignore_errors=True)


@pytest.mark.gpu
@pytest.mark.skipif(not (refrun.have_ref() and refrun.have_dropin()), reason="oracle/_ref harnesses not built")
@pytest.mark.parametrize("name,calc_paths", [("det_ties_small", 0), ("det_two_lane_congested", 0), ("route_choice_det", 0),
                                             ("det_gridlock", 0), ("rate_changes_det", 0), ("route_pruning_det", 0), ("signals_det", 0),
                                             ("demand_slices_det", 0), ("incidents_det", 0), ("vehicle_mix_det", 0),
                                             ("routes:all_from_sssp", 1),
                                             ("routes:partial_routes", 1)])
def test_mezzo_lib_with_dropin_step_gpu(built, tmp_path, name, calc_paths):
    """The drop-in proof for hot path 1: the reference's own executable flow (readmaster, executemaster, init, writeall —
    all unmodified mezzo_lib code) with ONLY the event loop of Network::step replaced by the GPU engine through the C++ host
    glue mezzo_b200/host/NetworkStep.h (oracle/_ref/mezzo_dropin_run). Every file Network::writeall derives from the two
    hot paths, the OD grids and the per-vehicle link exits must equal the all-CPU reference run.
    With calc_paths=1 the reference's own Network::shortest_paths_all runs too — compiled against the drop-in Graph class
    (mezzo_b200/host/Graph.h), so every labelCorrecting tree of the route generation is built on the GPU as well."""
    if name.startswith("routes:"):
        import test_routes
        spec = synth.sim_grid(**test_routes.CASES[name.split(":")[1]])
    else:
        spec = simcases.make(name)
    ref = os.path.join(str(tmp_path), "ref")
    mine = os.path.join(str(tmp_path), "dropin")
    a = refrun.run_reference(spec, workdir=ref, calc_paths=calc_paths, keep=True, save=True)
    b = refrun.run_reference(spec, workdir=mine, calc_paths=calc_paths, keep=True, save=True, binary=refrun.DROPIN_RUN)
    for fn in OUT_FILES:
        assert filecmp.cmp(os.path.join(ref, fn), os.path.join(mine, fn), shallow=False), fn
    assert np.array_equal(a["arrivals"], b["arrivals"])
    ea, eb = np.sort(a["exits"], order=["vid", "entry"]), np.sort(b["exits"], order=["vid", "entry"])
    assert len(ea) > 1000 and np.array_equal(ea, eb)
    assert a["info"]["currenttime"] == b["info"]["currenttime"]
