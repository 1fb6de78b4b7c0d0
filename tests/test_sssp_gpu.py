"""GPU parity tests for hot path 2 (route generation): the CUDA path, called through the C ABI, against the CPU
oracle (oracle/sssp_oracle.c) and the committed golden vectors of the unmodified reference.
Bar: BIT-EXACT predecessor arrays, node costs (IEEE doubles compared with ==) and route sets."""
import os

import numpy as np
import pytest

import cases
import mezzo_b200
import orc
from mezzo_b200 import _lib, synth

pytestmark = pytest.mark.gpu
MODES = [mezzo_b200.MZ_SSSP_EXACT, mezzo_b200.MZ_SSSP_AUTO]


def _gpu_graph(c):
    order = c["add_order"]
    g = mezzo_b200.Graph(c["n_nodes"], c["n_links"])
    for l in order:
        g.addLink(int(l), int(c["up"][l]), int(c["dn"][l]))
    g.set_downlink_indices()
    for a, b in c["prohibitors"]:
        g.set_turning_prohibitor(a, b)
    g.finalize()
    g.set_linktimes(c["T"], c["P"], c["periodlength"])
    return g


@pytest.fixture(scope="module")
def golden():
    return np.load(os.path.join(os.path.dirname(__file__), "golden", "sssp_golden.npz"))


@pytest.mark.parametrize("mode", MODES)
def test_golden_vectors_bit_exact(golden, mode):
    for trial in range(int(golden["n_cases"])):
        c = cases.make_case(trial)
        g = _gpu_graph(c)
        lp, npd, nc = g.sssp_batch(c["roots"], c["entries"], mode=mode)
        assert (lp == golden[f"c{trial}_link_pred"]).all(), trial
        assert (npd == golden[f"c{trial}_node_pred"]).all(), trial
        assert (nc == golden[f"c{trial}_node_cost"]).all(), trial
        # route sets for every (tree, node) pair
        n = len(c["roots"])
        N = c["n_nodes"]
        dest_off = np.arange(n + 1, dtype=np.int64) * N
        dests = np.tile(np.arange(N, dtype=np.int32), n)
        poff, links = g.sssp_routes(c["roots"], c["entries"], dest_off, dests, mode=mode)
        assert (poff == golden[f"c{trial}_routes_off"]).all(), trial
        assert (links == golden[f"c{trial}_routes"]).all(), trial
        g.close()


@pytest.mark.parametrize("mode", MODES)
def test_random_cases_vs_oracle(built, mode):
    for trial in range(500, 620):
        c = cases.make_case(trial)
        spec = orc.GraphSpec(c["n_nodes"], c["n_links"], c["up"], c["dn"], c["add_order"], c["prohibitors"])
        og = orc.OracleGraph(spec)
        og.set_linktimes(c["P"], c["periodlength"], c["T"])
        g = _gpu_graph(c)
        assert (g.legal.view(np.int8) == spec.legal).all()
        lp, npd, nc = g.sssp_batch(c["roots"], c["entries"], mode=mode)
        olp, onp, onc, _, canon = og.batch(c["roots"], c["entries"])
        assert (lp == olp).all() and (npd == onp).all() and (nc == onc).all(), trial
        st = g.last_stats()
        assert st["canonical_relax"] == canon, trial
        if mode == mezzo_b200.MZ_SSSP_EXACT:
            assert st["evaluated_relax"] >= canon
        g.close()


def _grid_case(nx, ny, P, seed, quantum=None, fifo=True):
    gl = synth.grid_links(nx, ny, seed=seed)
    T = synth.link_time_table(gl["length"], P, seed=seed, fifo=fifo, quantum=quantum)
    return gl, T


@pytest.mark.parametrize("mode", MODES)
@pytest.mark.parametrize("quantum,fifo", [(None, True), (10.0, True), (None, False)])
def test_grid_40x40_all_origins_vs_oracle(built, mode, quantum, fifo):
    """6 240-link grid, 256 trees over 4 entry times: continuous costs (tie-free), quantised costs (massive ties,
    order-dependent predecessors) and a non-FIFO table."""
    gl, T = _grid_case(40, 40, 4, 7, quantum, fifo)
    spec = orc.GraphSpec(gl["n_nodes"], gl["n_links"], gl["up"], gl["dn"])
    og = orc.OracleGraph(spec)
    og.set_linktimes(4, 900.0, T)
    g = mezzo_b200.Graph.from_arrays(gl["n_nodes"], gl["n_links"], gl["up"], gl["dn"])
    g.set_linktimes(T, 4, 900.0)
    rng = np.random.default_rng(3)
    roots = rng.choice(gl["n_links"], 256, replace=False).astype(np.int32)
    entries = (rng.integers(0, 4, 256) * 900.0).astype(np.float64)
    lp, npd, nc = g.sssp_batch(roots, entries, mode=mode)
    olp, onp, onc, _, canon = og.batch(roots, entries)
    assert (lp == olp).all()
    assert (npd == onp).all()
    assert (nc == onc).all()
    assert g.last_stats()["canonical_relax"] == canon
    g.close()


def test_chunked_batch_equals_single_chunk(built):
    """Work space is bounded: forcing tiny chunks (double-buffered copy-out) must not change a bit."""
    gl, T = _grid_case(12, 9, 3, 11)
    g = mezzo_b200.Graph.from_arrays(gl["n_nodes"], gl["n_links"], gl["up"], gl["dn"])
    g.set_linktimes(T, 3, 600.0)
    roots = np.arange(0, gl["n_links"], 3, dtype=np.int32)
    entries = np.linspace(0, 1700, len(roots))
    ref = g.sssp_batch(roots, entries)
    dest_off = np.arange(len(roots) + 1, dtype=np.int64) * 5
    dests = np.tile(np.array([0, 7, 50, 100, 107], np.int32), len(roots))
    r_ref = g.sssp_routes(roots, entries, dest_off, dests)
    g.set_option(_lib.MZ_OPT_CHUNK_TREES, 7)
    got = g.sssp_batch(roots, entries)
    assert g.last_stats()["n_chunks"] == (len(roots) + 6) // 7
    for a, b in zip(ref, got):
        assert (a == b).all()
    r_got = g.sssp_routes(roots, entries, dest_off, dests)
    assert (r_ref[0] == r_got[0]).all() and (r_ref[1] == r_got[1]).all()
    g.close()


def test_edge_cases():
    # single link, dest == its own down node; empty batch; unreachable destination; root validation
    g = mezzo_b200.Graph.from_arrays(3, 2, np.array([0, 2], np.int32), np.array([1, 0], np.int32))
    g.set_linktimes(np.array([[5.0], [np.nan]]), 1, 10.0)
    lp, npd, nc = g.sssp_batch([0], [0.0])
    assert lp.tolist() == [[-1, -2]] and npd.tolist() == [[-2, 0, -2]] and nc[0, 1] == 5.0
    lp, npd, nc = g.sssp_batch([1], [0.0])  # NaN period costs 0.1; link 0 reached through node 0
    assert lp.tolist() == [[1, -1]] and nc[0].tolist() == [0.1, 5.1, 9999999.0]
    lp, npd, nc = g.sssp_batch(np.empty(0, np.int32), np.empty(0))
    assert lp.shape == (0, 2)
    poff, links = g.sssp_routes([0, 0], [0.0, 0.0], [0, 2, 3], [1, 2, 1])
    assert poff.tolist() == [0, 1, 1, 2] and links.tolist() == [0, 0]
    with pytest.raises(mezzo_b200.MzError) as e:
        g.sssp_batch([5], [0.0])
    assert e.value.code == _lib.MZ_EINVAL
    g2 = mezzo_b200.Graph.from_arrays(3, 2, np.array([0, 2], np.int32), np.array([1, 0], np.int32))
    with pytest.raises(mezzo_b200.MzError) as e:
        g2.sssp_batch([0], [0.0])  # link times never set
    assert e.value.code == _lib.MZ_ESTATE
    g.close(), g2.close()


def test_device_resident_outputs(built):
    """Output pointers may be device memory (torch tensors): no staging copy, same bits."""
    import torch
    gl, T = _grid_case(10, 10, 2, 5)
    g = mezzo_b200.Graph.from_arrays(gl["n_nodes"], gl["n_links"], gl["up"], gl["dn"])
    g.set_linktimes(torch.from_numpy(T).cuda(), 2, 900.0)
    roots = np.arange(0, 64, dtype=np.int32)
    entries = np.zeros(64)
    ref = g.sssp_batch(roots, entries)
    out = (torch.empty((64, gl["n_links"]), dtype=torch.int32, device="cuda"),
           torch.empty((64, gl["n_nodes"]), dtype=torch.int32, device="cuda"),
           torch.empty((64, gl["n_nodes"]), dtype=torch.float64, device="cuda"))
    g.sssp_batch(roots, entries, out=out)
    torch.cuda.synchronize()
    for a, b in zip(ref, out):
        assert (a == b.cpu().numpy()).all()
    g.close()


def test_reference_style_single_tree_api(built):
    """labelCorrecting / reachable / shortest_path_vector used like Network::shortest_paths_all does."""
    c = cases.make_case(2)
    spec = orc.GraphSpec(c["n_nodes"], c["n_links"], c["up"], c["dn"], c["add_order"], c["prohibitors"])
    og = orc.OracleGraph(spec)
    og.set_linktimes(c["P"], c["periodlength"], c["T"])
    g = _gpu_graph(c)
    root = int(c["roots"][0])
    g.labelCorrecting(root, 0.0)
    lp, npd, nc, _, _ = og.tree(root, 0.0)
    for d in range(c["n_nodes"]):
        assert bool(g.reachable(d)) == (npd[d] != -2)
        if g.reachable(d):
            path = g.shortest_path_vector(d)
            if path and path[0] != root:
                path.insert(0, root)
            assert path == og.route(root, lp, npd, d).tolist()
    g.close()


@pytest.mark.skipif(not os.path.exists(os.path.join(orc.ORACLE_DIR, "_ref", "libdropin_sssp.so")),
                    reason="oracle/_ref/libdropin_sssp.so not built")
def test_cpp_dropin_graph_header(built):
    """mezzo_b200/host/Graph.h shadows the reference's Graph.h: the SAME C++ wrapper source that drives the reference
    (oracle/ref_sssp.cpp) is compiled against the drop-in class + the reference's own LinkTimeInfo and must give
    identical trees and routes — i.e. mezzo_lib code written against Graph<double,LinkTimeInfo> runs unchanged."""
    import ctypes as C
    lib = C.CDLL(os.path.join(orc.ORACLE_DIR, "_ref", "libdropin_sssp.so"))
    saved, orc._ref = orc._ref, None
    try:
        real = orc.ref_sssp_path
        orc.ref_sssp_path = lambda: os.path.join(orc.ORACLE_DIR, "_ref", "libdropin_sssp.so")
        for trial in (0, 1, 2, 5, 9, 13):
            c = cases.make_case(trial)
            spec = orc.GraphSpec(c["n_nodes"], c["n_links"], c["up"], c["dn"], c["add_order"], c["prohibitors"])
            dg = orc.RefGraph(spec)  # drives libdropin_sssp.so
            dg.set_linktimes(c["P"], c["periodlength"], c["T"], has_row=(c["up"] != -2).astype(np.int8))
            og = orc.OracleGraph(spec)
            og.set_linktimes(c["P"], c["periodlength"], c["T"])
            assert (dg.dn_legal() == spec.legal)[c["up"] != -2].all()
            for root, entry in zip(c["roots"], c["entries"]):
                lp, npd, nc, _ = dg.tree(int(root), float(entry))
                olp, onp, onc, _, _ = og.tree(int(root), float(entry))
                assert (lp == olp).all() and (npd == onp).all() and (nc == onc).all(), trial
                for d in range(0, c["n_nodes"], 2):
                    assert (dg.route(int(root), d) == og.route(int(root), olp, onp, d)).all()
    finally:
        orc.ref_sssp_path = real
        orc._ref = saved
    del lib


# ----------------------------------------------------------------------------------------------- batched kernel (AUTO)
def _auto_case(nx, ny, P, pl, seed, keep=1.0, n_trees=200, entries_lo_half=True):
    gl = synth.grid_links(nx, ny, seed=seed, keep_prob=keep, len_lo=50.0, len_hi=500.0)
    T = synth.link_time_table(gl["length"], P, seed=seed, fifo=True)
    rng = np.random.default_rng(seed)
    roots = rng.choice(gl["n_links"], n_trees).astype(np.int32)  # with repeats: several entry times per root
    hi = max(1, P // 2) if entries_lo_half else P
    entries = (rng.integers(0, hi, n_trees) * pl + rng.random(n_trees) * 0.5 * pl).astype(np.float64)
    return gl, T, roots, entries


@pytest.mark.parametrize("keep", [1.0, 0.625])
def test_auto_certifies_fifo_tie_free_trees(built, keep):
    """FIFO table long enough for every label, continuous costs: the batched (lanes = trees) kernel must certify every
    tree itself (no exact re-computation) and still be bit-exact against the oracle."""
    gl, T, roots, entries = _auto_case(30, 24, 6, 4000.0, 5, keep=keep)
    spec = orc.GraphSpec(gl["n_nodes"], gl["n_links"], gl["up"], gl["dn"])
    og = orc.OracleGraph(spec)
    og.set_linktimes(6, 4000.0, T)
    g = mezzo_b200.Graph.from_arrays(gl["n_nodes"], gl["n_links"], gl["up"], gl["dn"])
    g.set_linktimes(T, 6, 4000.0)
    lp, npd, nc = g.sssp_batch(roots, entries, mode=mezzo_b200.MZ_SSSP_AUTO)
    st = g.last_stats()
    olp, onp, onc, _, canon = og.batch(roots, entries)
    assert (lp == olp).all() and (npd == onp).all() and (nc == onc).all()
    assert st["n_exact_redo"] == 0, st
    assert st["fast_steps"] > 0
    assert st["canonical_relax"] == canon
    # bucket width is a tuning knob only
    for milli in (300, 1000, 20000):
        g.set_option(_lib.MZ_OPT_DELTA_MILLI, milli)
        lp2, np2, nc2 = g.sssp_batch(roots, entries, mode=mezzo_b200.MZ_SSSP_AUTO)
        assert (lp2 == olp).all() and (np2 == onp).all() and (nc2 == onc).all(), milli
        assert g.last_stats()["n_exact_redo"] == 0
    g.close()


def test_auto_mixed_certified_and_redone(built):
    """Late entry times run past the end of the table (cost 0.1 there: not FIFO) — those trees, and only those, are
    recomputed by the exact kernel; the batch as a whole stays bit-exact. A step cap forces the abort path."""
    gl, T, roots, entries = _auto_case(25, 25, 4, 1500.0, 9, entries_lo_half=False, n_trees=300)
    spec = orc.GraphSpec(gl["n_nodes"], gl["n_links"], gl["up"], gl["dn"])
    og = orc.OracleGraph(spec)
    og.set_linktimes(4, 1500.0, T)
    g = mezzo_b200.Graph.from_arrays(gl["n_nodes"], gl["n_links"], gl["up"], gl["dn"])
    g.set_linktimes(T, 4, 1500.0)
    olp, onp, onc, _, canon = og.batch(roots, entries)
    lp, npd, nc = g.sssp_batch(roots, entries, mode=mezzo_b200.MZ_SSSP_AUTO)
    st = g.last_stats()
    assert (lp == olp).all() and (npd == onp).all() and (nc == onc).all()
    assert 0 < st["n_exact_redo"] < len(roots), st
    g.set_option(_lib.MZ_OPT_MAX_STEPS, 5)
    lp, npd, nc = g.sssp_batch(roots, entries, mode=mezzo_b200.MZ_SSSP_AUTO)
    assert (lp == olp).all() and (npd == onp).all() and (nc == onc).all()
    assert g.last_stats()["n_exact_redo"] == len(roots)
    g.close()


def test_auto_many_bundles_and_chunks(built):
    """More than one bundle per chunk, several chunks, a ragged last bundle, device-resident outputs."""
    torch = pytest.importorskip("torch")
    gl, T, roots, entries = _auto_case(16, 16, 3, 6000.0, 21, n_trees=333)
    spec = orc.GraphSpec(gl["n_nodes"], gl["n_links"], gl["up"], gl["dn"])
    og = orc.OracleGraph(spec)
    og.set_linktimes(3, 6000.0, T)
    olp, onp, onc, _, _ = og.batch(roots, entries)
    g = mezzo_b200.Graph.from_arrays(gl["n_nodes"], gl["n_links"], gl["up"], gl["dn"])
    g.set_linktimes(T, 3, 6000.0)
    for chunk in (0, 100, 32, 7):
        g.set_option(_lib.MZ_OPT_CHUNK_TREES, chunk)
        lp, npd, nc = g.sssp_batch(roots, entries, mode=mezzo_b200.MZ_SSSP_AUTO)
        assert (lp == olp).all() and (npd == onp).all() and (nc == onc).all(), chunk
        assert g.last_stats()["n_exact_redo"] == 0
    g.set_option(_lib.MZ_OPT_CHUNK_TREES, 0)
    dev = torch.device("cuda", 0)
    out = (torch.empty((len(roots), gl["n_links"]), dtype=torch.int32, device=dev),
           torch.empty((len(roots), gl["n_nodes"]), dtype=torch.int32, device=dev),
           torch.empty((len(roots), gl["n_nodes"]), dtype=torch.float64, device=dev))
    g.sssp_batch(roots, entries, mode=mezzo_b200.MZ_SSSP_AUTO, out=out)
    torch.cuda.synchronize()
    assert (out[0].cpu().numpy() == olp).all() and (out[1].cpu().numpy() == onp).all()
    assert (out[2].cpu().numpy() == onc).all()
    g.close()
