"""Generates tests/golden/sssp_golden.npz by running the UNMODIFIED reference (oracle/_ref/libref_sssp.so =
B/Graph.h + B/Graph.cpp + B/linktimes.cpp compiled in place) on the seeded cases of tests/cases.py.
Run in the authoring container only (needs /root/reference to have built oracle/_ref):

    make -C oracle ref && python tests/golden/make_sssp_golden.py
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
import cases  # noqa: E402
import orc  # noqa: E402

N_CASES = 60


def main():
    out = {}
    for trial in range(N_CASES):
        c = cases.make_case(trial)
        spec = orc.GraphSpec(c["n_nodes"], c["n_links"], c["up"], c["dn"], c["add_order"], c["prohibitors"])
        rg = orc.RefGraph(spec)
        rg.set_linktimes(c["P"], c["periodlength"], c["T"], has_row=(c["up"] != -2).astype(np.int8))
        lps, nps, ncs, ncalls, routes_off, routes = [], [], [], [], [0], []
        for root, entry in zip(c["roots"], c["entries"]):
            lp, npd, nc, calls = rg.tree(int(root), float(entry))
            lps.append(lp), nps.append(npd), ncs.append(nc), ncalls.append(calls)
            for d in range(c["n_nodes"]):
                r = rg.route(int(root), d)
                routes.extend(r.tolist())
                routes_off.append(len(routes))
        out[f"c{trial}_legal"] = rg.dn_legal()
        out[f"c{trial}_link_pred"] = np.array(lps, np.int32)
        out[f"c{trial}_node_pred"] = np.array(nps, np.int32)
        out[f"c{trial}_node_cost"] = np.array(ncs, np.float64)
        out[f"c{trial}_ncost_calls"] = np.array(ncalls, np.int64)
        out[f"c{trial}_routes_off"] = np.array(routes_off, np.int64)
        out[f"c{trial}_routes"] = np.array(routes, np.int32)
    out["n_cases"] = np.array(N_CASES)
    np.savez_compressed(os.path.join(HERE, "sssp_golden.npz"), **out)
    print("wrote", os.path.join(HERE, "sssp_golden.npz"))


if __name__ == "__main__":
    main()
