"""Seeded SSSP test cases shared by the golden generator, the oracle tests and the GPU parity tests."""
import numpy as np

from mezzo_b200 import synth

SP_UNSET = -2


def make_case(trial):
    """Returns dict(spec args, P, periodlength, T, roots, entries) for a small graph; covers: random graphs with
    out-degree > 8 (8-bit dnLegal truncation), grids with integer costs (ties), thinned grids (unreachable parts),
    sparse link ids (unused slots), shuffled addLink order, turn prohibitors, NaN periods, non-FIFO tables,
    entry times beyond the table."""
    rng = np.random.default_rng(1000 + trial)
    kind = trial % 3
    if kind == 0:
        g = synth.random_links(int(rng.integers(5, 60)), int(rng.integers(5, 300)), seed=trial,
                               max_out=int(rng.integers(2, 14)))
    elif kind == 1:
        g = synth.grid_links(int(rng.integers(2, 9)), int(rng.integers(2, 9)), seed=trial, int_lengths=bool(trial & 1))
    else:
        g = synth.grid_links(int(rng.integers(3, 9)), int(rng.integers(3, 9)), seed=trial, keep_prob=0.7)
    L, N = g["n_links"], g["n_nodes"]
    slots = L + int(rng.integers(0, 5))
    ids = np.sort(rng.choice(slots, L, replace=False))
    up = np.full(slots, SP_UNSET, np.int32)
    dn = np.full(slots, SP_UNSET, np.int32)
    up[ids] = g["up"]
    dn[ids] = g["dn"]
    order = ids.copy()
    if trial % 5 == 0:
        rng.shuffle(order)
    pro = []
    for _ in range(int(rng.integers(0, 6))):
        a = int(rng.choice(ids))
        cands = ids[up[ids] == dn[a]]
        if len(cands):
            pro.append((a, int(rng.choice(cands))))
    pro = list(dict.fromkeys(pro))  # a repeated prohibitor trips the reference's assert (B/Graph.cpp:716)
    P = int(rng.integers(1, 6))
    pl = float(rng.choice([30.0, 100.0, 900.0]))
    length = np.zeros(slots)
    length[ids] = g["length"]
    T = synth.link_time_table(length, P, seed=trial, fifo=bool(trial & 2),
                              missing_prob=0.1 if trial % 4 == 0 else 0.0, quantum=10.0 if trial % 3 == 1 else None)
    T[up == SP_UNSET] = np.nan
    k = min(4, L)
    roots = rng.choice(ids, k, replace=False).astype(np.int32)
    entries = rng.choice([0.0, 17.5, pl * P * 0.9, pl * P * 2], k).astype(np.float64)
    return dict(n_nodes=N, n_links=slots, up=up, dn=dn, add_order=order.astype(np.int32), prohibitors=pro, P=P,
                periodlength=pl, T=T, roots=roots, entries=entries)
