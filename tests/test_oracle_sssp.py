"""CPU tests (no GPU): the C restatement oracle/sssp_oracle.c against
  (1) the committed golden vectors produced by the unmodified reference (tests/golden/sssp_golden.npz), and
  (2) the compiled reference itself (oracle/_ref/libref_sssp.so) when it is present."""
import numpy as np
import pytest

import cases
import orc


@pytest.fixture(scope="module")
def golden():
    import os
    return np.load(os.path.join(os.path.dirname(__file__), "golden", "sssp_golden.npz"))


def _oracle_for(c):
    spec = orc.GraphSpec(c["n_nodes"], c["n_links"], c["up"], c["dn"], c["add_order"], c["prohibitors"])
    og = orc.OracleGraph(spec)
    og.set_linktimes(c["P"], c["periodlength"], c["T"])
    return spec, og


def test_oracle_matches_golden(built, golden):
    n = int(golden["n_cases"])
    assert n >= 60
    for trial in range(n):
        c = cases.make_case(trial)
        spec, og = _oracle_for(c)
        assert (spec.legal == golden[f"c{trial}_legal"])[c["up"] != -2].all()
        roff = golden[f"c{trial}_routes_off"]
        routes = golden[f"c{trial}_routes"]
        k = 0
        for i, (root, entry) in enumerate(zip(c["roots"], c["entries"])):
            lp, npd, nc, ncalls, canon = og.tree(int(root), float(entry))
            assert (lp == golden[f"c{trial}_link_pred"][i]).all(), (trial, i)
            assert (npd == golden[f"c{trial}_node_pred"][i]).all(), (trial, i)
            assert (nc == golden[f"c{trial}_node_cost"][i]).all(), (trial, i)  # bit-exact doubles
            assert ncalls == golden[f"c{trial}_ncost_calls"][i]
            assert canon <= ncalls - 1
            for d in range(c["n_nodes"]):
                r = og.route(int(root), lp, npd, d)
                assert (r == routes[roff[k]:roff[k + 1]]).all(), (trial, i, d)
                k += 1


@pytest.mark.skipif(orc.ref_lib() is None, reason="oracle/_ref not built (reference tree absent)")
def test_oracle_matches_compiled_reference(built):
    for trial in range(100, 400):
        c = cases.make_case(trial)
        spec, og = _oracle_for(c)
        rg = orc.RefGraph(spec)
        rg.set_linktimes(c["P"], c["periodlength"], c["T"], has_row=(c["up"] != -2).astype(np.int8))
        for root, entry in zip(c["roots"], c["entries"]):
            a = og.tree(int(root), float(entry))
            b = rg.tree(int(root), float(entry))
            assert (a[0] == b[0]).all() and (a[1] == b[1]).all() and (a[2] == b[2]).all() and a[3] == b[3], trial
            for d in range(0, c["n_nodes"], 3):
                assert (og.route(int(root), a[0], a[1], d) == rg.route(int(root), d)).all()


def test_oracle_known_answers(built):
    """Hand-checked tiny cases of the quirks listed in SURVEY.md H2 / A10."""
    # chain 0->1->2->3 plus a link back into the root's up-node (must be skipped, B/Graph.cpp:385)
    up = np.array([0, 1, 2, 2], np.int32)
    dn = np.array([1, 2, 3, 0], np.int32)
    spec = orc.GraphSpec(4, 4, up, dn)
    og = orc.OracleGraph(spec)
    T = np.array([[1.0], [2.0], [4.0], [8.0]])
    og.set_linktimes(1, 100.0, T)
    lp, npd, nc, ncalls, canon = og.tree(0, 0.0)
    assert lp.tolist() == [-1, 0, 1, -2]
    assert npd.tolist() == [-2, 0, 1, 2]
    assert nc.tolist() == [9999999.0, 1.0, 3.0, 7.0]
    assert canon == 2 and ncalls == 3
    # beyond the table / absent period costs 0.1 (B/linktimes.cpp:74-80)
    lp, npd, nc, _, _ = og.tree(0, 1000.0)
    assert nc[1] == 0.1 and abs(nc[3] - 0.3) < 1e-12
    # a node with 10 out-links: the 9th and 10th (dnIndex >= 8) are never relaxed (char truncation, B/Graph.h:127)
    up = np.array([0] + [1] * 10, np.int32)
    dn = np.array([1] + list(range(2, 12)), np.int32)
    og2 = orc.OracleGraph(orc.GraphSpec(12, 11, up, dn))
    og2.set_linktimes(1, 100.0, np.ones((11, 1)))
    lp, npd, _, _, canon = og2.tree(0, 0.0)
    assert lp[1:9].tolist() == [0] * 8 and lp[9:].tolist() == [-2, -2]
    assert canon == 8
    # turn prohibitor 0->2 removes that relaxation only
    spec3 = orc.GraphSpec(12, 11, up, dn, prohibitors=[(0, 2)])
    og3 = orc.OracleGraph(spec3)
    og3.set_linktimes(1, 100.0, np.ones((11, 1)))
    lp, _, _, _, _ = og3.tree(0, 0.0)
    assert lp[1] == 0 and lp[2] == -2 and lp[3] == 0
    # equal labels: strict '<' keeps the first predecessor found (B/Graph.cpp:412)
    up = np.array([0, 1, 1, 2, 3], np.int32)  # 0:0->1, 1:1->2, 2:1->3, 3:2->4, 4:3->4 ... then 5: 4->5
    dn = np.array([1, 2, 3, 4, 4], np.int32)
    up = np.append(up, 4).astype(np.int32)
    dn = np.append(dn, 5).astype(np.int32)
    og4 = orc.OracleGraph(orc.GraphSpec(6, 6, up, dn))
    og4.set_linktimes(1, 100.0, np.ones((6, 1)))
    lp, npd, nc, _, _ = og4.tree(0, 0.0)
    assert lp[5] == 3 and npd[4] == 3 and nc[5] == 4.0
