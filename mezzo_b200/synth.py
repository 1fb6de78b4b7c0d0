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


# --------------------------------------------------------------------------------------------------------
# Link-update networks (hot path 1): a NetSpec is a plain dict in the REFERENCE's own vocabulary and ids
# (net.dat / turnings.dat / demand.dat / routes.dat / histtimes.dat / vehiclemix.dat).
# --------------------------------------------------------------------------------------------------------
def sim_grid(nx, ny, seed=0, od_every=5, n_od=None, trips_per_hour=None, total_trips=None, horizon=3600.0,
             lanes=1, len_lo=300, len_hi=700, connector_len=250, server_type=1, server_mu=1.8, server_sd=0.3,
             shared_server=False, vfree=15.0, vmin=2.0, romax=150.0, romin=10.0, lookback=20, veh_length=6.0,
             standard_veh_length=7, randseed=42, n_periods=4, dest_server=(1, 0.5, 0.1, 0.0), two_lane_prob=0.0,
             sd_type=0, sd_alpha=1.0, sd_beta=1.0, alt_routes=1, hist_variation=0.0, od_servers_deterministic=1, keep_routes=1.0, hist_fifo=False,
             od_radius=None, drop_prob=0.0, rate_changes=0, param_server_type=1, server_delay=0.0, signals=0, delete_bad_routes=0, max_rel_route_cost=1.5, small_od_rate=0.0, demand_slices=0, incidents=0, vehicle_classes=None):
    """nx×ny junction grid, bidirectional integer-length links, an origin and a destination (with connectors)
    at every `od_every`-th junction, explicit turnings for every non-U-turn movement, one server per turning
    (sid = tid) unless shared_server, one route per OD pair (shortest by free-flow time),
    deterministic OD servers, a single vehicle class. Config 2 = sim_grid(100, 100, total_trips=200_000).
    `od_radius` (in OD-lattice steps) is the recipe for networks too large for an all-sources Dijkstra: every pair's
    destination lies within that many lattice steps of its origin and its route is a random monotone staircase
    through the grid (routes.dat routes need not be shortest paths — the reference drives whatever the file holds).
    `drop_prob` deletes grid links at random (config 3's irregular graph); with od_radius it needs keep_routes=0, i.e.
    every route then comes from the route generation (calc_paths=1)."""
    from scipy.sparse import csr_matrix
    from scipy.sparse.csgraph import dijkstra

    rng = np.random.default_rng(seed)
    J = nx * ny
    jid = lambda k: 1 + k  # noqa: E731  junction ids 1..J
    nodes = [(jid(k), 3, float((k % nx) * 100), float((k // nx) * 100), None) for k in range(J)]
    links = []  # (lid, in, out, length, lanes, sdid)
    lid = 1
    for k in range(J):
        x, y = k % nx, k // nx
        for dx, dy in ((1, 0), (-1, 0), (0, 1), (0, -1)):
            xx, yy = x + dx, y + dy
            if 0 <= xx < nx and 0 <= yy < ny:
                if drop_prob > 0 and rng.random() < drop_prob:
                    continue  # an irregular network: this direction of the street does not exist
                nl = 2 if rng.random() < two_lane_prob else lanes
                links.append((lid, jid(k), jid(yy * nx + xx), int(rng.integers(len_lo, len_hi + 1)), nl, 0))
                lid += 1
    od_j = [k for k in range(J) if (k % nx) % od_every == od_every // 2 and (k // nx) % od_every == od_every // 2]
    if not od_j:
        od_j = [0, J - 1]
    origins, dests = [], []
    servers = []  # (sid, type, mu, sd, delay)
    next_node = J + 1
    sid_dest0 = 1_000_000 if 16 * J < 1_000_000 else 1_000_000_000  # above every turning id (sid = tid)
    for i, k in enumerate(od_j):
        o, d = next_node, next_node + 1
        next_node += 2
        nodes.append((o, 1, 0.0, 0.0, None))
        dsid = sid_dest0 + i
        servers.append((dsid, dest_server[0], dest_server[1], dest_server[2], dest_server[3]))
        nodes.append((d, 2, 0.0, 0.0, dsid))
        links.append((lid, o, jid(k), int(connector_len), lanes, 0))
        lid += 1
        links.append((lid, jid(k), d, int(connector_len), lanes, 0))
        lid += 1
        origins.append(o)
        dests.append(d)
    links_a = np.array(links, dtype=np.int64)
    # turnings: every (in-link, out-link) at a junction except the U-turn
    turnings = []
    by_in = {}
    for r in links_a:
        by_in.setdefault(int(r[1]), []).append(r)
    tid = 1
    if shared_server:
        servers.append((0, server_type, server_mu, server_sd, float(server_delay)))
    for r in links_a:
        n = int(r[2])
        if n > J:  # destination
            continue
        for o in by_in.get(n, []):
            if int(o[2]) == int(r[1]):
                continue  # U-turn
            sid = 0 if shared_server else tid
            if not shared_server:
                servers.append((sid, server_type, server_mu, server_sd, float(server_delay)))
            turnings.append((tid, n, sid, int(r[0]), int(o[0]), lookback))
            tid += 1
    # routes: one per OD pair, shortest by free-flow time on the node graph
    idx_of_node = {n[0]: i for i, n in enumerate(nodes)}
    N = len(nodes)
    w = links_a[:, 3] / vfree
    rows = np.array([idx_of_node[int(a)] for a in links_a[:, 1]])
    cols = np.array([idx_of_node[int(a)] for a in links_a[:, 2]])
    # keep the cheapest parallel link only (there are none in a grid, but be safe)
    G = csr_matrix((w, (rows, cols)), shape=(N, N))
    link_of = {(int(a), int(b)): int(l) for l, a, b in zip(links_a[:, 0], rows, cols)}
    n_pairs = n_od if n_od is not None else min(4000, len(origins) * (len(dests) - 1))
    pairs = set()
    if od_radius is not None:
        return _finish_local(locals())
    while len(pairs) < n_pairs and len(pairs) < len(origins) * (len(dests) - 1):
        a, b = int(rng.integers(len(origins))), int(rng.integers(len(dests)))
        if a != b:
            pairs.add((a, b))
    pairs = sorted(pairs)
    src = sorted({a for a, _ in pairs})
    _, pred = dijkstra(G, directed=True, indices=[idx_of_node[origins[a]] for a in src], return_predecessors=True)
    src_row = {a: i for i, a in enumerate(src)}
    routes, ods, alt_jobs = [], [], []
    if total_trips is not None:
        rate = total_trips / max(1, len(pairs)) * 3600.0 / horizon
    else:
        rate = trips_per_hour if trips_per_hour is not None else 60.0
    rid = 1
    for a, b in pairs:
        o, d = origins[a], dests[b]
        path, cur = [], idx_of_node[d]
        prow = pred[src_row[a]]
        while cur != idx_of_node[o]:
            p = int(prow[cur])
            if p < 0:
                path = None
                break
            path.append(link_of[(p, cur)])
            cur = p
        if not path:
            continue
        routes.append((rid, o, d, path[::-1]))
        ods.append((o, d, float(rate)))
        rid += 1
        alt_jobs.append((a, b, tuple(path[::-1])))
    # alternative routes (route choice): shortest paths under randomly perturbed link weights
    if alt_routes > 1:
        seen = {}
        for a, b, base in alt_jobs:
            seen[(a, b)] = {base}
        for k in range(1, alt_routes):
            wk = w * (1.0 + 1.5 * rng.random(len(w)))
            Gk = csr_matrix((wk, (rows, cols)), shape=(N, N))
            _, predk = dijkstra(Gk, directed=True, indices=[idx_of_node[origins[a]] for a in src], return_predecessors=True)
            for a, b, _ in alt_jobs:
                o, d = origins[a], dests[b]
                path, cur = [], idx_of_node[d]
                prow = predk[src_row[a]]
                while cur != idx_of_node[o]:
                    p = int(prow[cur])
                    if p < 0:
                        path = None
                        break
                    path.append(link_of[(p, cur)])
                    cur = p
                if not path:
                    continue
                path = tuple(path[::-1])
                if path in seen[(a, b)]:
                    continue
                seen[(a, b)].add(path)
                routes.append((rid, o, d, list(path)))
                rid += 1
    return _finish_spec(locals())


def _finish_spec(v):
    """Common tail of sim_grid: the NetSpec dict from the generator's local state."""
    g = lambda k: v[k]  # noqa: E731
    rng, links_a, routes = g("rng"), g("links_a"), g("routes")
    sd_type, vfree, n_periods, horizon = g("sd_type"), g("vfree"), g("n_periods"), g("horizon")
    sdf = (0, sd_type, vfree, g("vmin"), g("romax"), g("romin")) + ((g("sd_alpha"), g("sd_beta")) if sd_type else ())
    if g("keep_routes") < 1.0:
        # route generation (calc_paths=1): only a part of the OD pairs comes with routes, the rest is found by the SSSP
        keep = rng.random(len(routes)) < g("keep_routes")
        # ids 1..n: shortest_paths_all numbers new routes from routemap.size()+1 and asserts that the id is free
        routes = [(i + 1,) + tuple(r[1:]) for i, r in enumerate(r for r, k in zip(routes, keep) if k)]
    histtimes = None
    if g("hist_variation") > 0:
        # time-varying historical link times (histtimes.dat): what Route::cost integrates for the route choice
        ff = np.maximum(1.0, links_a[:, 3] / vfree)
        ht = ff[:, None] * (1.0 + g("hist_variation") * rng.random((len(links_a), int(n_periods))))
        if g("hist_fifo"):
            ht = np.sort(ht, axis=1)  # non-decreasing over the periods: a FIFO table (the frontier SSSP kernel certifies it)
        histtimes = {int(l): [float(x) for x in row] for l, row in zip(links_a[:, 0], ht)}
    serverrates = []
    if g("rate_changes") > 0:
        # serverrates.dat: turning servers change their mean and/or sd at some point of the horizon;
        # a value of -1 leaves that parameter alone, a few servers change twice, two changes share one time instant
        turn_sids = [t[2] for t in g("turnings")]
        pick = rng.choice(len(turn_sids), size=min(int(g("rate_changes")), len(turn_sids)), replace=False)
        for j, k in enumerate(pick.tolist()):
            tm = float(np.round(rng.uniform(0.1, 0.8) * horizon, 1))
            mu = float(np.round(g("server_mu") * rng.uniform(0.6, 2.5), 2))
            serverrates.append((turn_sids[k], tm, mu if j % 3 != 1 else -1.0, 0.2 if j % 3 != 0 else -1.0))
            if j % 4 == 0:
                serverrates.append((turn_sids[k], tm if j % 8 == 0 else float(np.round(tm + 0.1 * horizon, 1)),
                                    float(np.round(g("server_mu") * 1.3, 2)), -1.0))
    sig = []
    if g("signals") > 0:
        # signal.dat: one control per chosen junction, two plans (the second takes over halfway and ends before the horizon,
        # after which its turnings stay red), one stage per in-link of the junction (all turnings leaving that link)
        by_node = {}
        for t in g("turnings"):
            by_node.setdefault(t[1], {}).setdefault(t[3], []).append(t[0])
        cand = sorted(nd for nd, m in by_node.items() if len(m) >= 2)
        pick = rng.choice(len(cand), size=min(int(g("signals")), len(cand)), replace=False)
        sid = 1
        for ci, c in enumerate(sorted(pick.tolist())):
            groups = [by_node[cand[c]][l] for l in sorted(by_node[cand[c]])]
            plans = []
            for pi in range(2):
                stages = []
                for grp in groups:
                    stages.append((sid, 0.0, float(rng.integers(15, 40)) + (0.5 if ci % 2 else 0.0), list(grp)))
                    sid += 1
                start = 0.0 if pi == 0 else float(np.round(0.5 * horizon))
                stop = float(np.round(0.55 * horizon)) if pi == 0 else float(np.round(0.9 * horizon))
                plans.append((10 * (ci + 1) + pi, start, stop, 0.0, 90.0, stages))
            sig.append((ci + 1, plans))
    slices, ods = [], g("ods")
    if g("demand_slices") > 0:
        # demand.dat slices (row f1: MatrixAction): a third of the OD pairs changes its rate at each load time — up, down, to
        # zero (the pair stops for good) — and a few pairs start inactive (rate 0 or below 1) until a slice switches them on
        ods = list(ods)
        for i in range(0, len(ods), 7):
            ods[i] = (ods[i][0], ods[i][1], 0.0 if i % 2 else 0.5)
        K = int(g("demand_slices"))
        for j in range(K):
            lt = float(np.round((j + 1) / (K + 1) * horizon, 1))
            pick = np.sort(rng.choice(len(ods), size=max(1, len(ods) // 3), replace=False))
            base = g("ods")
            rates = [(base[i][0], base[i][1], int(base[i][2] * float(rng.choice([0.0, 0.5, 1.7, 2.5])))) for i in pick.tolist()]
            slices.append((lt, 1.0 if j % 2 == 0 else 1.5, rates))
    incident_sdfuncs, incident_list = [], []
    if g("incidents") > 0:
        # incident file (row f4, no information broadcast): a slow speed-density function replaces the link's own between
        # start and stop on the busiest links of the route set; two incidents overlap on one link (Link::temp_sdfunc then
        # restores the FIRST incident's function for good, B/link.cpp:854-868)
        incident_sdfuncs = [(900, 0, 0.35 * vfree, g("vmin"), g("romax"), g("romin")), (901, 0, 0.6 * vfree, 1.0, 120.0, 5.0)]
        use = {}
        for r in routes:
            for l in r[3][1:-1]:
                use[l] = use.get(l, 0) + 1
        busy = [l for l, _ in sorted(use.items(), key=lambda x: (-x[1], x[0]))][:int(g("incidents"))]
        for j, l in enumerate(busy):
            start = float(np.round((0.15 + 0.5 * rng.random()) * horizon, 1))
            stop = float(np.round(start + (0.1 + 0.2 * rng.random()) * horizon, 1))
            incident_list.append((int(l), 900 + j % 2, 0.0, start, stop, -1.0, -1.0, j % 2))
        l0, s0, e0 = incident_list[0][0], incident_list[0][3], incident_list[0][4]
        incident_list.append((l0, 901, 0.0, float(np.round(0.5 * (s0 + e0), 1)), float(np.round(e0 + 0.05 * horizon, 1)), -1.0, -1.0, 0))
    return dict(
        incident_sdfuncs=incident_sdfuncs, incidents=incident_list,
        slices=slices, signals=sig, serverrates=serverrates, servers=g("servers"), nodes=g("nodes"), sdfuncs=[sdf],
        links=[tuple(r) for r in links_a.tolist()], turnings=g("turnings"), ods=ods, routes=routes,
        # vehicle_classes = [(probability, length)]: several classes in vehiclemix.dat, drawn per trip (B/od.cpp:79)
        vtypes=([(i + 1, f"class{i + 1}", float(p), float(ln)) for i, (p, ln) in enumerate(g("vehicle_classes"))] if g("vehicle_classes")
                else [(1, "cars", 1.0, float(g("veh_length")))]), n_periods=int(n_periods),
        periodlength=float(horizon) / n_periods, histtimes=histtimes,  # None ⇒ "links: 0" ⇒ free-flow times
        params=dict(standard_veh_length=int(g("standard_veh_length")), server_type=int(g("param_server_type")),
                    od_servers_deterministic=int(g("od_servers_deterministic")),
                    update_interval_routes=900.0, default_lookback_size=20, use_giveway=0, kirchoff_alpha=-1.0,
                    delete_bad_routes=int(g("delete_bad_routes")), max_rel_route_cost=float(g("max_rel_route_cost")),
                    small_od_rate=float(g("small_od_rate"))),
        starttime=0.0, stoptime=float(horizon), randseed=int(g("randseed")))


def _finish_local(v):
    """od_radius mode of sim_grid: local OD pairs, staircase routes (no shortest-path search)."""
    rng, nx, od_every, od_j, origins, dests = v["rng"], v["nx"], v["od_every"], v["od_j"], v["origins"], v["dests"]
    n_pairs, R, link_of, idx_of_node = v["n_pairs"], int(v["od_radius"]), v["link_of"], v["idx_of_node"]
    cx = sorted({k % nx for k in od_j})
    cy = sorted({k // nx for k in od_j})
    ncx, ncy = len(cx), len(cy)
    at = {(k % nx, k // nx): i for i, k in enumerate(od_j)}
    pairs = set()
    cap = len(origins) * min(len(dests) - 1, (2 * R + 1) ** 2 - 1)
    while len(pairs) < min(n_pairs, cap):
        m = max(1024, n_pairs - len(pairs))
        a = rng.integers(len(origins), size=m)
        ax, ay = a % ncx, a // ncx  # od_j is row-major over the OD lattice
        bx = np.clip(ax + rng.integers(-R, R + 1, size=m), 0, ncx - 1)
        by = np.clip(ay + rng.integers(-R, R + 1, size=m), 0, ncy - 1)
        for i, j in zip(a.tolist(), (by * ncx + bx).tolist()):
            if i != j and len(pairs) < n_pairs:
                pairs.add((i, j))
    pairs = sorted(pairs)
    assert all(at[(cx[i % ncx], cy[i // ncx])] == i for i in (0, len(od_j) - 1))
    total_trips, horizon = v["total_trips"], v["horizon"]
    if total_trips is not None:
        rate = total_trips / max(1, len(pairs)) * 3600.0 / horizon
    else:
        rate = v["trips_per_hour"] if v["trips_per_hour"] is not None else 60.0
    J = nx * v["ny"]
    n_links_grid = len(v["links_a"]) - 2 * len(od_j)
    lid_conn0 = n_links_grid + 1  # connectors were appended pairwise: origin i -> lid_conn0 + 2 i, dest i -> +1
    routes, ods = [], []
    flips = rng.random(len(pairs))
    if v["keep_routes"] <= 0.0:  # no routes.dat at all: only the demand
        ods = [(origins[a], dests[b], float(rate)) for a, b in pairs]
        pairs = []
    assert not (pairs and v["drop_prob"] > 0), "staircase routes need the full grid: use keep_routes=0 with drop_prob"
    for rid0, (a, b) in enumerate(pairs):
        ka, kb = od_j[a], od_j[b]
        x, y, xb, yb = ka % nx, ka // nx, kb % nx, kb // nx
        path = [lid_conn0 + 2 * a]
        # staircase: alternate runs in x and y of random length until the destination junction is reached
        horiz = flips[rid0] < 0.5
        step = 1 + int(flips[rid0] * 1e6) % (od_every + 2)
        while (x, y) != (xb, yb):
            for _ in range(step):
                if horiz and x != xb:
                    nxt = (x + (1 if xb > x else -1), y)
                elif not horiz and y != yb:
                    nxt = (x, y + (1 if yb > y else -1))
                else:
                    break
                path.append(link_of[(y * nx + x, nxt[1] * nx + nxt[0])])
                x, y = nxt
            horiz = not horiz
        path.append(lid_conn0 + 2 * b + 1)
        routes.append((rid0 + 1, origins[a], dests[b], path))
        ods.append((origins[a], dests[b], float(rate)))
    v = dict(v)
    v["routes"], v["ods"] = routes, ods
    return _finish_spec(v)
