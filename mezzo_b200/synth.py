"""Synthetic road graphs and link-time tables for the configs named in BASELINE.json.

All generators are seeded and numpy-only; ids are dense so they can be used as array slots
(the reference indexes its SSSP arrays by raw id, B/network.cpp:6598).
"""
import numpy as np

SP_UNSET = -2


def grid_links(nx, ny, seed=0, len_lo=300.0, len_hi=700.0, keep_prob=1.0, int_lengths=False):
    """nx×ny junction grid with bidirectional links between 4-neighbours.

    keep_prob < 1 thins directed links independently (config 4: 1000×1000 thinned to out-degree ≈2.5
    ⇒ keep_prob = 0.625 of the ≈4 per junction). Returns dict with up, dn (int32 [L]), length (float64 [L]),
    x, y (float64 [N]). Link ids are dense in generation order (ascending by up-node, E/W/N/S)."""
    rng = np.random.default_rng(seed)
    idx = np.arange(nx * ny, dtype=np.int64).reshape(ny, nx)
    pairs = []
    pairs.append((idx[:, :-1].ravel(), idx[:, 1:].ravel()))  # east
    pairs.append((idx[:, 1:].ravel(), idx[:, :-1].ravel()))  # west
    pairs.append((idx[:-1, :].ravel(), idx[1:, :].ravel()))  # north
    pairs.append((idx[1:, :].ravel(), idx[:-1, :].ravel()))  # south
    up = np.concatenate([p[0] for p in pairs])
    dn = np.concatenate([p[1] for p in pairs])
    if keep_prob < 1.0:
        keep = rng.random(len(up)) < keep_prob
        up, dn = up[keep], dn[keep]
    order = np.lexsort((dn, up))  # ids ascending by up node: spatial locality of ids
    up, dn = up[order], dn[order]
    length = rng.uniform(len_lo, len_hi, len(up))
    if int_lengths:
        length = np.floor(length)
    ys, xs = np.divmod(np.arange(nx * ny), nx)
    return dict(n_nodes=nx * ny, n_links=len(up), up=up.astype(np.int32), dn=dn.astype(np.int32),
                length=length, x=xs.astype(np.float64) * 500.0, y=ys.astype(np.float64) * 500.0)


def random_links(n_nodes, n_links, seed=0, max_out=8, len_lo=50.0, len_hi=500.0):
    """Random directed multigraph with out-degree capped at max_out (the reference silently ignores the
    9th+ out-link of a node, SURVEY.md H2 — pass max_out>8 to exercise that quirk)."""
    rng = np.random.default_rng(seed)
    up = rng.integers(0, n_nodes, n_links * 2)
    dn = rng.integers(0, n_nodes, n_links * 2)
    cnt = np.zeros(n_nodes, np.int64)
    keep = np.zeros(len(up), bool)
    n = 0
    for i in range(len(up)):
        if cnt[up[i]] < max_out:
            cnt[up[i]] += 1
            keep[i] = True
            n += 1
            if n == n_links:
                break
    up, dn = up[keep], dn[keep]
    return dict(n_nodes=n_nodes, n_links=len(up), up=up.astype(np.int32), dn=dn.astype(np.int32),
                length=rng.uniform(len_lo, len_hi, len(up)))


def link_time_table(length, n_periods, seed=0, speed=13.9, spread=0.5, fifo=True, missing_prob=0.0,
                    quantum=None):
    """T[l][p] = length/speed · (1 + spread·u)  (SURVEY.md §8d cfg 4).

    fifo=True makes each row non-decreasing in p (piece-wise constant costs are FIFO iff non-decreasing);
    quantum rounds costs to multiples of `quantum` to manufacture exact ties;
    missing_prob blanks entries with NaN (period absent ⇒ cost 0.1, B/linktimes.cpp:74-80)."""
    rng = np.random.default_rng(seed + 7919)
    L = len(length)
    T = (np.asarray(length)[:, None] / speed) * (1.0 + spread * rng.random((L, n_periods)))
    if fifo:
        T = np.maximum.accumulate(T, axis=1)
    if quantum:
        T = np.maximum(quantum, np.round(T / quantum) * quantum)
    if missing_prob > 0:
        T[rng.random(T.shape) < missing_prob] = np.nan
    return np.ascontiguousarray(T)
