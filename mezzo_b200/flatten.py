"""Host-side flattening layer for hot path 1: NetSpec (the reference's own vocabulary: servers/nodes/sdfuncs/links/
turnings/od pairs/routes, reference ids) → MzNet (dense structure-of-arrays, include/mezzo_b200.h), including the
host pre-pass that generates the trip table in the reference's ODaction order.

Reference behaviour restated here (B/ = /root/reference/branches/busmezzo/mezzo_lib/src/):
  Network::init order and init times           B/network.cpp:7657-7731
  ODpair::execute first booking                B/od.cpp:355-368    (per-pair Random seeded with randseed, B/od.cpp:319)
  ODaction::execute / ODServer::next           B/od.cpp:69-108, B/server.cpp:86-91
  Random::urandom                              B/Random.cpp:85-98
This is index bookkeeping and a few LCG draws per OD pair — no traffic arithmetic happens on the host."""
import ctypes as C

import math

import numpy as np

from ._lib import MzNetStruct


class Lcg:
    """Random (Park–Miller A=48271, M=2^31-1), B/Random.cpp:85-98."""
    M, A = 2147483647, 48271
    Q, R = M // A, M % A

    def __init__(self, seed):
        self.seed = int(seed)

    def urandom(self):
        s = self.A * (self.seed % self.Q) - self.R * (self.seed // self.Q)
        # C++ '%' and '/' truncate toward zero; seeds stay positive after the first draw so floor == trunc here
        self.seed = s if s > 0 else s + self.M
        return float(self.seed) / float(self.M)

    def urandom_ab(self, a, b):
        return a + (b - a) * self.urandom()


class RouteCost:
    """Route::cost (B/route.cpp:173-197) over LinkTime::cost (B/linktimes.cpp:36-82), including its inverted cache
    test: the cost is recomputed on every call until a call comes more than update_interval_routes after the last
    computation — from then on the stale sum is returned forever."""

    def __init__(self, hist_rows, periodlength, interval):
        self.rows, self.pl, self.interval = hist_rows, float(periodlength), float(interval)
        self.sumcost, self.last = 0.0, 0.0

    def cost(self, time):
        if self.sumcost > 0.0 and self.last + self.interval < time:
            return self.sumcost
        temp = time
        self.last = time
        for row in self.rows:
            i = int(temp / self.pl)  # static_cast<int>, then unsigned: negative times cannot occur
            temp += float(row[i]) if 0 <= i < len(row) else 0.1
        self.sumcost = temp - time
        return self.sumcost


def select_route(costs, alpha, rnd):
    """ODpair::select_route (B/od.cpp:247-300): Kirchhoff shares pow(cost, kirchoff_alpha), one urandom draw."""
    import math
    util = [math.pow(c, alpha) for c in costs]
    total = 0.0
    for u in util:
        total += u
    choice = rnd.urandom()
    sub = 0.0
    for i, u in enumerate(util):
        sub += u
        if sub / total >= choice:
            return i
    return 0  # "TROUBLE: No Route selected! Returning the first in list"


class MzDemandStruct(C.Structure):
    """MzDemand of include/mezzo_b200.h."""
    _fields_ = [("n_odpairs", C.c_int32), ("n_routes", C.c_int32), ("n_periods", C.c_int32), ("n_turnings", C.c_int32),
                ("od_rate", C.c_void_p), ("od_route_off", C.c_void_p), ("od_routes", C.c_void_p), ("route_off", C.c_void_p),
                ("route_links", C.c_void_p), ("link_hist", C.c_void_p), ("periodlength", C.c_double), ("runtime", C.c_double),
                ("update_interval_routes", C.c_double), ("kirchoff_alpha", C.c_double),
                ("od_servers_deterministic", C.c_int32), ("n_signal_controls", C.c_int32), ("randseed", C.c_int64),
                ("route_sumcost0", C.c_void_p), ("route_last0", C.c_void_p),
                ("n_slice_rates", C.c_int32), ("slice_od", C.c_void_p), ("slice_time", C.c_void_p), ("slice_rate", C.c_void_p),
                ("phantom_final", C.c_void_p), ("n_transit_init_steps", C.c_int32)]


def slice_rates(spec, od_sorted):
    """demand.dat slices (Network::readrates / readrate, B/network.cpp:5509-5606) → the rate changes MatrixAction::execute
    applies (B/network.cpp:8213-8229), flattened to (OD index into od_sorted, load time, rate) in execution order: by load
    time, equal times in file order (the MatrixActions were booked while the file was read). readrate reads the rate as an
    int and truncates rate*scale to an int again."""
    idx = {}
    for k, (o, d, _r) in enumerate(od_sorted):
        idx.setdefault((o, d), k)  # find_if over the sorted OD pairs: the first pair with these ids
    out = []
    for loadtime, scale, rates in sorted(spec.get("slices") or [], key=lambda x: x[0]):
        for o, d, rate in rates:
            if (o, d) not in idx:
                raise ValueError(f"demand slice at {loadtime}: OD pair ({o}, {d}) is not in the OD list")
            r = int(int(rate) * float(scale))  # static_cast<int>(rate*scale) on an int rate
            if r < 0:
                raise ValueError(f"demand slice at {loadtime}: negative rate for OD pair ({o}, {d})")
            out.append((idx[(o, d)], float(loadtime), float(r)))
    return out


def prune_routes(spec, od_sorted, od_route_idx, route_hist):
    """ODpair::delete_spurious_routes(runtime/4) for every OD pair, as Network::add_od_routes runs it when
    delete_bad_routes=1 (B/network.cpp:5426-5441, B/od.cpp:160-244): routes costing more than threshold x the cheapest are
    dropped, the rest is sorted by cost(0.0) and cut to maxroutes. Returns (od_route_idx pruned and re-ordered, and the
    (sumcost, last_calc_time) every surviving Route carries into the run — the pricing goes through Route::cost)."""
    P = spec["params"]
    small, max_rel = float(P.get("small_od_rate", 3.0)), float(P.get("max_rel_route_cost", 2.0))
    interval, pl = float(P.get("update_interval_routes", 900.0)), float(spec["periodlength"])
    time = spec["stoptime"] / 4
    n_routes = len(route_hist)
    sum0, last0 = np.zeros(n_routes), np.zeros(n_routes)
    out = []
    for (o, d, rate), routes in zip(od_sorted, od_route_idx):
        if not routes:
            out.append([])
            continue
        rc = {r: RouteCost(route_hist[r], pl, interval) for r in routes}
        min_cost = rc[routes[0]].cost(time)
        for r in routes:
            if rc[r].cost(time) < min_cost:
                min_cost = rc[r].cost(time)
        if rate < small:
            threshold, maxroutes = 1.5, 5
        else:
            q = rate / small if small != 0 else math.inf
            # static_cast<int>(r+1): out of range / inf gives INT_MIN on x86-64, read back as unsigned 2^31 = no limit
            maxroutes = int(q + 1) if math.isfinite(q) and q + 1 < 2 ** 31 else 2 ** 31
            threshold = max_rel
        keep = [r for r in routes if not rc[r].cost(time) > threshold * min_cost]
        # sort(routes, compare_route_cost): cost(0.0) on both sides of every comparison; insertion sort below 16 elements
        # in libstdc++, i.e. stable — equal costs keep their order
        import functools
        keep.sort(key=functools.cmp_to_key(lambda a, b: -1 if rc[a].cost(0.0) < rc[b].cost(0.0) else
                                           (1 if rc[b].cost(0.0) < rc[a].cost(0.0) else 0)))
        for r in keep:
            rc[r].cost(0.0)  # the "cost of sorted routes" print loop prices every survivor at time 0
        del keep[maxroutes:]
        for r in keep:
            sum0[r], last0[r] = rc[r].sumcost, rc[r].last
        out.append(keep)
    return out, sum0, last0


def vtype_probabilities(vtypes):
    """Vtypes::initialize (B/vtypes.cpp:52-64): the class probabilities divided by their sum unless that is exactly 1. The
    reference then sorts the POINTERS (`sort(vtypes.begin(), vtypes.end())` over vector<Vtype*>, B/vtypes.cpp:66), i.e. by heap
    address — for the objects Network::readvtypes allocates back to back that is the file order, which is the order used here
    (pinned against the compiled reference by tests/test_trips.py::test_vehicle_classes_match_the_reference)."""
    p = [float(v[2]) for v in vtypes]
    total = 0.0
    for x in p:
        total += x
    return p if total == 1 else [x / total for x in p]


def trip_lengths(vtypes, randseed, n_trips, native=True):
    """Vehicle length of every trip: ODaction::execute draws the vehicle class from ONE network-wide Random stream
    (`odpair->vehtypes()->random_vtype()`, B/od.cpp:79, B/vtypes.h:61-71), one draw per trip in global event order.
    `vtypes` = [(id, label, probability, length)] of vehiclemix.dat, or a single length. Returns (length[n], class index[n])."""
    if not isinstance(vtypes, (list, tuple)):
        return np.full(n_trips, float(vtypes)), np.zeros(n_trips, np.int32)
    if len(vtypes) == 1:  # (the draw still happens in the reference, but nothing else reads that stream)
        return np.full(n_trips, float(vtypes[0][3])), np.zeros(n_trips, np.int32)
    prob = np.ascontiguousarray(vtype_probabilities(vtypes), np.float64)
    idx = np.zeros(max(1, n_trips), np.int32)
    if native:
        from ._lib import check, lib
        l = lib()
        l.mz_vtypes_draw.argtypes = [C.c_int64, C.c_int32, C.c_void_p, C.c_int64, C.c_void_p]
        check(l.mz_vtypes_draw(int(randseed), len(prob), prob.ctypes.data_as(C.c_void_p), int(n_trips),
                               idx.ctypes.data_as(C.c_void_p)))
    else:
        rnd = Lcg(randseed)
        for i in range(n_trips):
            r, acc, pick = rnd.urandom(), 0.0, len(prob) - 1
            for k in range(len(prob)):
                acc += prob[k]
                if acc >= r:
                    pick = k
                    break
            idx[i] = pick
    idx = idx[:n_trips]
    return np.array([float(v[3]) for v in vtypes])[idx], idx


def generate_trips(spec, n_turnings, od_sorted, od_route_idx, origin_index, veh_length, route_off, route_links, link_hist,
                   route_state=None):
    """The trip table through the native host pre-pass mz_trips_generate (mezzo_b200/csrc/trips.cu); same return value as
    generate_trips_py, which restates the same reference code in Python and is what the tests compare it with."""
    from ._lib import check, lib
    l = lib()
    l.mz_trips_generate.argtypes = [C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
                                    C.POINTER(C.c_int64), C.POINTER(C.c_int32)]
    a = lambda v, dt: np.ascontiguousarray(v, dt)  # noqa: E731
    rate = a([x[2] for x in od_sorted], np.float64)
    cnt = np.array([len(r) for r in od_route_idx], np.int64)
    od_off = a(np.concatenate([[0], np.cumsum(cnt)]), np.int32)
    od_routes = a([r for rs in od_route_idx for r in rs], np.int32)
    route_off, route_links, link_hist = a(route_off, np.int32), a(route_links, np.int32), a(link_hist, np.float64)
    d = MzDemandStruct()
    d.n_odpairs, d.n_routes, d.n_periods, d.n_turnings = len(od_sorted), len(route_off) - 1, link_hist.shape[1], n_turnings
    ptr = lambda x: x.ctypes.data_as(C.c_void_p)  # noqa: E731
    d.od_rate, d.od_route_off, d.od_routes = ptr(rate), ptr(od_off), ptr(od_routes)
    d.route_off, d.route_links, d.link_hist = ptr(route_off), ptr(route_links), ptr(link_hist)
    d.periodlength, d.runtime = float(spec["periodlength"]), float(spec["stoptime"])
    d.update_interval_routes = float(spec["params"].get("update_interval_routes", 900.0))
    d.kirchoff_alpha = float(spec["params"].get("kirchoff_alpha", -1.0))
    d.od_servers_deterministic = int(bool(spec["params"].get("od_servers_deterministic", 1)))
    d.randseed = int(spec["randseed"])
    d.n_signal_controls = len(spec.get("signals") or [])
    d.n_transit_init_steps = int(spec.get("n_transit_init_steps", 0))  # a BusMezzo scenario's buslines / ODstops pairs
    if route_state is not None:
        sum0, last0 = a(route_state[0], np.float64), a(route_state[1], np.float64)
        d.route_sumcost0, d.route_last0 = ptr(sum0), ptr(last0)
    sl = slice_rates(spec, od_sorted)
    if sl:
        sl_od, sl_t, sl_r = a([x[0] for x in sl], np.int32), a([x[1] for x in sl], np.float64), a([x[2] for x in sl], np.float64)
        d.n_slice_rates, d.slice_od, d.slice_time, d.slice_rate = len(sl), ptr(sl_od), ptr(sl_t), ptr(sl_r)
    phantom = np.zeros(2)
    d.phantom_final = ptr(phantom)
    n, n_init = C.c_int64(0), C.c_int32(0)
    check(l.mz_trips_generate(C.byref(d), 0, None, None, None, None, C.byref(n), C.byref(n_init)))
    t, tprev = np.empty(max(1, n.value)), np.empty(max(1, n.value))
    od, route = np.empty(max(1, n.value), np.int32), np.empty(max(1, n.value), np.int32)
    check(l.mz_trips_generate(C.byref(d), n.value, ptr(t), ptr(tprev), ptr(od), ptr(route), C.byref(n), C.byref(n_init)))
    t, tprev, od, route = t[:n.value], tprev[:n.value], od[:n.value], route[:n.value]
    init_idx = (np.cumsum(cnt > 0) - 1).astype(np.int32)  # index among the initialised OD pairs
    origin = np.array([origin_index[x[0]] for x in od_sorted], np.int32)
    return (t, origin[od] if len(od) else od, route, trip_lengths(veh_length, spec["randseed"], len(t))[0], od, int(n_init.value),
            init_idx[od] if len(od) else od, tprev, (float(phantom[0]), float(phantom[1])))


def generate_trips_py(spec, n_turnings, od_sorted, od_route_idx, origin_index, veh_length, route_hist=None,
                      route_state=None):
    """ODpair::execute + ODaction::execute for every OD pair, merged in global event order, with the rate changes of the
    demand slices (MatrixAction::execute → ODpair::set_rate → ODaction::set_rate, B/network.cpp:8213-8229, B/od.h:63,102).
    Returns (trip_time, trip_origin_idx, trip_route_idx, trip_length, trip_od_idx (into od_sorted), n_odpairs_initialised,
    trip_od_init_idx (index among the initialised OD pairs = MzNet.trip_od), trip_prev_time, phantom_final).
    phantom_final = (t, t_parent) of the earliest event at/after the horizon that generates nothing — a MatrixAction or the
    poll of an inactive ODaction (B/od.cpp:102-107) — or (0, 0): if it precedes every other event beyond the horizon it is
    the one event Network::step still executes there, so no action of the network runs."""
    runtime = spec["stoptime"]
    det = bool(spec["params"].get("od_servers_deterministic", 1))
    interval = float(spec["params"].get("update_interval_routes", 900.0))
    initvalue = 0.1
    for _ in range(n_turnings):
        initvalue += 0.00001  # one per Turning::init call, accumulated like the reference does
    for _ in spec.get("signals") or []:
        initvalue += 0.000001  # one per SignalControl (B/network.cpp:7666-7670)
    for _ in range(int(spec.get("n_transit_init_steps", 0))):
        initvalue += 0.00001  # one per Busline and per active ODstops pair (B/network.cpp:7672-7694)
    by_od = {}
    phantom = None
    for k, lt, r in slice_rates(spec, od_sorted):
        by_od.setdefault(k, []).append((lt, r))
    for lt, _scale, _rates in spec.get("slices") or []:
        # the MatrixAction itself: booked while the demand file was read, i.e. before every other event of its instant
        if lt >= runtime and (phantom is None or (lt, -math.inf) < phantom):
            phantom = (float(lt), -math.inf)
    times, ods, prevs, chosen = [], [], [], []
    n_init = 0
    for k, (o, d, rate) in enumerate(od_sorted):
        routes = od_route_idx[k]
        if not routes:
            continue  # "chuck out the OD pairs without paths": no init slot is consumed
        # the pair's own Random is seeded 42 whatever the run's seed (_DETERMINISTIC_ROUTE_CHOICE, B/od.cpp:313-320);
        # the ODaction's ODServer owns a second Random seeded with randseed (B/server.cpp:6-13)
        rnd = Lcg(42)
        srv = Lcg(spec["randseed"])
        t_init = initvalue
        initvalue += 0.00001
        k_init = n_init  # index among the initialised OD pairs
        n_init += 1
        active = rate >= 1  # ODaction::ODaction (B/od.cpp:30-46)
        mu = 3600.0 / rate if active else 0.0
        if rate > 0:  # ODpair::execute (B/od.cpp:355-368)
            t = t_init + rnd.urandom_ab(1.2, 3600 / rate)
        else:  # ... executes the (inactive) ODaction on the spot: it books a poll
            t = t_init + interval + rnd.urandom_ab(1, 100)
        sl, si = by_od.get(k, []), 0
        rc = None
        if len(routes) > 1:
            rc = [RouteCost(route_hist[r], spec["periodlength"], interval) for r in routes]
            if route_state is not None:
                for c, r in zip(rc, routes):
                    c.sumcost, c.last = float(route_state[0][r]), float(route_state[1][r])
        alpha = float(spec["params"].get("kirchoff_alpha", -1.0))
        prev = t_init  # the booking event: ODpair::execute, then the pair's previous event
        ts, ps, ch = [], [], []
        while math.isfinite(t):  # (a rate of 0 books the next event at +inf: the pair is dead for good)
            while si < len(sl) and sl[si][0] <= t:  # MatrixActions up to and including this instant come first
                mu = 3600.0 / sl[si][1] if sl[si][1] != 0 else math.inf
                active = True
                si += 1
            if not active:  # B/od.cpp:102-107: no vehicle, poll again after update_interval_routes + U(1, 100)
                if t >= runtime:
                    if phantom is None or (t, prev) < phantom:
                        phantom = (t, prev)
                    break
                prev, t = t, t + interval + rnd.urandom_ab(1, 100)
                continue
            # ODaction::execute → select_route(time): no draw with a single route; otherwise one draw of the pair's own
            # Random (the one that already gave the start offset), shares from Route::cost at the event's time. The pair's
            # events are the only callers of its routes' cost(), so evaluating them pair by pair keeps the cache state exact.
            ch.append(routes[0] if rc is None else routes[select_route([c.cost(t) for c in rc], alpha, rnd)])
            ts.append(t)
            ps.append(prev)
            if t >= runtime:
                break  # the first event at/after the horizon may still execute (B/network.cpp:7804-7806)
            # ODServer::next (B/server.cpp:82-87): time + mu, or time + rrandom(mu) = -log(urandom) * mu (B/Random.cpp:284)
            prev, t = t, (t + mu if det else t + (-math.log(srv.urandom()) * mu))
        if not ts:
            continue
        times.append(np.array(ts))
        prevs.append(np.array(ps))
        ods.append(np.stack([np.full(len(ts), k, np.int32), np.full(len(ts), k_init, np.int32)], 1))
        chosen.append(np.array(ch, np.int32))
    ph = phantom if phantom is not None else (0.0, 0.0)
    if not times:
        z = np.empty(0)
        return z, z.astype(np.int32), z.astype(np.int32), z, z.astype(np.int32), n_init, z.astype(np.int32), z, ph
    t = np.concatenate(times)
    tprev = np.concatenate(prevs)
    od2 = np.concatenate(ods)
    route = np.concatenate(chosen)
    order = np.lexsort((tprev, t))  # by time; equal times FIFO = by the booking event's time, then OD init order (stable)
    t, tprev, od2, route = t[order], tprev[order], od2[order], route[order]
    od, od_init = od2[:, 0], od2[:, 1]
    origin = np.array([origin_index[od_sorted[k][0]] for k in range(len(od_sorted))], np.int32)[od]
    return (t, origin, route, trip_lengths(veh_length, spec["randseed"], len(t), native=False)[0], od.astype(np.int32), n_init,
            od_init.astype(np.int32), tprev, ph)


def shared_stochastic_servers(spec):
    """{server id: sorted node ids} for every stochastic server (file type 1 or 4: one Random per Server object,
    B/server.cpp:9-13) that turnings or destinations of SEVERAL nodes name. Draw k of such a server is f^k(seed) and k is the
    number of draws made before it ANYWHERE in the network, so its stream follows the global event order: the
    engine refuses these networks (MZ_ELIMIT in mz_sim_create; DESIGN.md §7 has the argument) and this is the host-side
    check that names the offending servers. Servers shared inside one node are fine."""
    typ = {s[0]: s[1] for s in spec["servers"]}
    users = {}
    for t in spec["turnings"]:
        if t[2] >= 0 and typ.get(t[2]) in (1, 4):
            users.setdefault(t[2], set()).add(t[1])
    for n in spec["nodes"]:
        if n[1] == 2 and n[4] is not None and n[4] >= 0 and typ.get(n[4]) in (1, 4):
            users.setdefault(n[4], set()).add(n[0])
    return {sid: sorted(nodes) for sid, nodes in users.items() if len(nodes) > 1}


class FlatNet:
    """Dense arrays + id maps; .struct is the MzNet handed to mz_sim_create / the oracle."""

    def __init__(self, spec):
        self.spec = spec
        nodes = sorted(spec["nodes"], key=lambda n: n[0])
        links = sorted(spec["links"], key=lambda l: l[0])  # ascending id = linkmap order
        self.node_ids = np.array([n[0] for n in nodes], np.int64)
        self.link_ids = np.array([l[0] for l in links], np.int64)
        nidx = {int(n): i for i, n in enumerate(self.node_ids)}
        lidx = {int(l): i for i, l in enumerate(self.link_ids)}
        self.node_index, self.link_index = nidx, lidx
        # the incident file's own sdfuncs join the network's sdfunc map (Network::readincidentfile → readsdfuncs)
        sdf = sorted(list(spec["sdfuncs"]) + list(spec.get("incident_sdfuncs") or []), key=lambda s: s[0])
        sdidx = {int(s[0]): i for i, s in enumerate(sdf)}
        srv = sorted(spec["servers"], key=lambda s: s[0])
        sidx = {int(s[0]): i for i, s in enumerate(srv)}
        a = lambda v, dt: np.ascontiguousarray(v, dt)  # noqa: E731
        self.link_length = a([l[3] for l in links], np.int32)
        self.link_lanes = a([l[4] for l in links], np.int32)
        self.link_sdfunc = a([sdidx[l[5]] for l in links], np.int32)
        self.link_in_node = a([nidx[l[1]] for l in links], np.int32)
        self.link_out_node = a([nidx[l[2]] for l in links], np.int32)
        self.sd_type = a([s[1] for s in sdf], np.int32)
        self.sd_vfree = a([s[2] for s in sdf], np.float64)
        self.sd_vmin = a([s[3] for s in sdf], np.float64)
        self.sd_romax = a([s[4] for s in sdf], np.float64)
        self.sd_romin = a([s[5] for s in sdf], np.float64)
        self.sd_alpha = a([s[6] if len(s) > 6 else 0.0 for s in sdf], np.float64)
        self.sd_beta = a([s[7] if len(s) > 7 else 0.0 for s in sdf], np.float64)
        self.srv_type = a([s[1] for s in srv], np.int32)
        self.srv_mu = a([s[2] for s in srv], np.float64)
        self.srv_sd = a([s[3] for s in srv], np.float64)
        self.srv_delay = a([s[4] for s in srv], np.float64)
        turns = sorted([t for t in spec["turnings"] if t[2] >= 0], key=lambda t: t[0])  # sid<0 = SSSP prohibitor only
        self.turn_ids = np.array([t[0] for t in turns], np.int64)
        self.turn_node = a([nidx[t[1]] for t in turns], np.int32)
        self.turn_server = a([sidx[t[2]] for t in turns], np.int32)
        self.turn_in = a([lidx[t[3]] for t in turns], np.int32)
        self.turn_out = a([lidx[t[4]] for t in turns], np.int32)
        self.turn_lookback = a([t[5] for t in turns], np.int32)
        self.shared_stochastic_servers = shared_stochastic_servers(spec)
        dests = [n for n in nodes if n[1] == 2]
        origins = [n for n in nodes if n[1] == 1]
        self.dest_node = a([nidx[n[0]] for n in dests], np.int32)
        self.dest_server = a([sidx[n[4]] if n[4] is not None and n[4] >= 0 else -1 for n in dests], np.int32)
        self.origin_node = a([nidx[n[0]] for n in origins], np.int32)
        origin_index = {n[0]: i for i, n in enumerate(origins)}
        # free-flow / hist times (B/network.cpp:6254-6273: "links: 0" ⇒ every period = freeflow time)
        P = int(spec["n_periods"])
        vfree0 = np.array([self._speed0(sdf[i]) for i in self.link_sdfunc])
        ff = np.maximum(1.0, self.link_length / vfree0)
        if spec.get("histtimes") is None:
            self.link_hist = np.repeat(ff[:, None], P, axis=1).copy()
        else:
            self.link_hist = np.array([spec["histtimes"][int(l)] for l in self.link_ids], np.float64).reshape(len(links), P)
        self.freeflow = ff
        # routes (file order) and OD pairs (sorted by origin id, destination id — od_less_than)
        self.route_ids = np.array([r[0] for r in spec["routes"]], np.int64)
        off, rl = [0], []
        for r in spec["routes"]:
            rl.extend(lidx[l] for l in r[3])
            off.append(len(rl))
        self.route_off = a(off, np.int32)
        self.route_links = a(rl, np.int32)
        od_sorted = sorted(spec["ods"], key=lambda x: (x[0], x[1]))
        self.od_sorted = od_sorted
        by_od = {}
        for i, r in enumerate(spec["routes"]):
            by_od.setdefault((r[1], r[2]), []).append(i)
        od_route_idx = [by_od.get((o, d), []) for o, d, _ in od_sorted]
        self.route_state = None
        if spec["params"].get("delete_bad_routes", 0):
            hist_rows = [self.link_hist[self.route_links[self.route_off[r]:self.route_off[r + 1]]]
                         for r in range(len(self.route_ids))]
            od_route_idx, sum0, last0 = prune_routes(spec, od_sorted, od_route_idx, hist_rows)
            self.route_state = (sum0, last0)
        # routes that are still in some OD pair's choice set (all of them unless delete_bad_routes pruned the sets)
        self.route_kept = np.zeros(len(self.route_ids), bool)
        self.route_kept[[r for rs in od_route_idx for r in rs]] = True
        (self.trip_time, self.trip_origin, self.trip_route, self.trip_length, self.trip_od_sorted,
         self.n_odpairs, self.trip_od, self.trip_prev_time, self.phantom_final) = generate_trips(
             spec, len(turns), od_sorted, od_route_idx, origin_index, list(spec["vtypes"]),
             self.route_off, self.route_links, self.link_hist, route_state=self.route_state)
        self._trip_args = (len(turns), od_sorted, od_route_idx, origin_index)  # for the tests' Python cross-check
        self.trip_time = a(self.trip_time, np.float64)
        self.trip_origin = a(self.trip_origin, np.int32)
        self.trip_route = a(self.trip_route, np.int32)
        self.trip_length = a(self.trip_length, np.float64)
        self.trip_od = a(self.trip_od, np.int32)
        self.trip_prev_time = a(self.trip_prev_time, np.float64)
        self.link_hist = a(self.link_hist, np.float64)
        # serverrates.dat: (server id, time, mu, sd) → ChangeRateAction (B/network.cpp:5636-5665); file order is kept
        rates = spec.get("serverrates") or []
        self.chg_server = a([sidx[r[0]] for r in rates], np.int32)
        self.chg_time = a([r[1] for r in rates], np.float64)
        self.chg_mu = a([r[2] for r in rates], np.float64)
        self.chg_sd = a([r[3] for r in rates], np.float64)
        # signal.dat: [(control id, [(plan id, start, stop, offset, cycletime, [(stage id, start, duration, [turning ids])])])]
        tidx = {int(t): i for i, t in enumerate(self.turn_ids)}
        c_off, p_start, p_stop, p_off, s_dur, s_off, s_turns = [0], [], [], [0], [], [0], []
        for _cid, plans in spec.get("signals") or []:
            for _pid, start, stop, _offset, _cycle, stages in plans:
                p_start.append(start)
                p_stop.append(stop)
                for _sid, _sstart, duration, turns_ in stages:
                    s_dur.append(duration)
                    s_turns.extend(tidx[t] for t in turns_)
                    s_off.append(len(s_turns))
                p_off.append(len(s_dur))
            c_off.append(len(p_start))
        self.sigc_plan_off, self.sigp_stage_off, self.sigs_turn_off = a(c_off, np.int32), a(p_off, np.int32), a(s_off, np.int32)
        self.sigp_start, self.sigp_stop, self.sigs_duration = a(p_start, np.float64), a(p_stop, np.float64), a(s_dur, np.float64)
        self.sigs_turns = a(s_turns, np.int32)
        # incidents: (link id, sdfunc id, penalty, start, stop, info_start, info_stop, blocked), B/network.cpp:6337-6380
        inc = spec.get("incidents") or []
        for x in inc:
            if x[5] >= 0.0 and x[6] >= 0.0:
                raise NotImplementedError("incidents with information broadcast (en-route switching) are not built")
        self.inc_link = a([lidx[x[0]] for x in inc], np.int32)
        self.inc_sdfunc = a([sdidx[x[1]] for x in inc], np.int32)
        self.inc_start = a([x[3] for x in inc], np.float64)
        self.inc_stop = a([x[4] for x in inc], np.float64)
        self.struct = self._make_struct()

    def set_stops(self, trip_stop_off, stop_link, stop_pos, n_transit_init_steps=0):
        """Transit overlay hooks: the stops of every trip in visit order — CSR offsets per trip, dense
        link index and Busstop::position (m) per stop. A trip with stops is a type-4 vehicle (B/link.cpp:489-516)."""
        self.trip_stop_off = np.ascontiguousarray(trip_stop_off, np.int32)
        self.stop_link = np.ascontiguousarray(stop_link, np.int32)
        self.stop_pos = np.ascontiguousarray(stop_pos, np.float64)
        self.n_transit_init_steps = int(n_transit_init_steps)  # MzNet::n_transit_init_steps (B/network.cpp:7672-7694)
        if len(self.trip_stop_off) != len(self.trip_time) + 1 or self.trip_stop_off[-1] != len(self.stop_link):
            raise ValueError("trip_stop_off must have n_trips + 1 entries and end at the number of stops")
        self.struct = self._make_struct()
        return self

    @staticmethod
    def _speed0(sd):
        return sd[2]  # Sdfunc::speed(0): ro <= romin (romin >= 0 asserted by the reader) ⇒ vfree

    def _make_struct(self):
        s = MzNetStruct()
        sp = self.spec
        s.n_nodes, s.n_links = len(self.node_ids), len(self.link_ids)
        s.n_sdfuncs, s.n_servers = len(self.sd_type), len(self.srv_type)
        s.n_turnings, s.n_dests, s.n_origins = len(self.turn_node), len(self.dest_node), len(self.origin_node)
        s.n_routes, s.n_trips, s.n_periods = len(self.route_ids), len(self.trip_time), int(sp["n_periods"])
        s.n_odpairs = int(self.n_odpairs)
        s.n_rate_changes = len(self.chg_server)
        s.n_incidents = len(self.inc_link)
        s.n_sig_controls, s.n_sig_plans = len(self.sigc_plan_off) - 1, len(self.sigp_start)
        s.n_sig_stages = len(self.sigs_duration)
        s.standard_veh_length = int(sp["params"]["standard_veh_length"])
        s.server_type = int(sp["params"]["server_type"])
        s.use_giveway = int(sp["params"].get("use_giveway", 0))
        s.randseed = int(sp["randseed"])
        s.runtime = float(sp["stoptime"])
        s.periodlength = float(sp["periodlength"])
        s.phantom_final_t, s.phantom_final_tp = self.phantom_final
        s.n_transit_init_steps = int(getattr(self, "n_transit_init_steps", sp.get("n_transit_init_steps", 0)))
        for name, _ in MzNetStruct._fields_:
            if hasattr(self, name) and isinstance(getattr(self, name), np.ndarray):
                setattr(s, name, getattr(self, name).ctypes.data_as(C.c_void_p))
        return s
