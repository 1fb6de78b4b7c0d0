"""The reference's on-disk formats either side of the two hot paths (grammars §5f), so an existing
Mezzo scenario directory can be run through the engine and leaves the files existing Mezzo tooling reads.

B/ = /root/reference/branches/busmezzo/mezzo_lib/src/. Restated here:
  readers   Network::readmaster B/network.cpp:6954-7190, Parameters::read_parameters B/parameters.cpp:118+,
            readservers/readnodes/readsdfuncs/readlinks 494-804, readturnings 1039-1091, readdemandfile 5337-5424,
            readlinktimes/readtimes 6204-6324, readroutes 1204-1304, readvtypes 5795-5850
  writers   Network::writelinktimes 6189-6202 + Link::write_time B/link.cpp:741-752, Network::writepathfile 6505-6517 +
            Route::write B/route.cpp:220-228, Network::writeoutput / writesummary 5854-5883 + Grid::write_empty B/grid.cpp:17-37
  driver    Network::executemaster 7296-7392 (read order, calc_paths, routes.dat written back) → Network::step → writeall 7393-7420

The text writers use the stream's default formatting (6 significant digits) like the reference's `out << double`.
Everything outside the two hot paths that a scenario may switch on (incidents WITH information broadcast, virtual links,
boundary nodes, give-ways, transit) is REFUSED with NotImplementedError, never ignored.
"""
import os
import re

import numpy as np

MASTER_KEYS = ["network=", "turnings=", "signals=", "histtimes=", "routes=", "demand=", "incident=", "vehicletypes=",
               "virtuallinks=", "serverrates=", "#output_files", "linktimes=", "output=", "summary=", "speeds=", "inflows=",
               "outflows=", "queuelengths=", "densities=", "#scenario", "starttime=", "stoptime=", "calc_paths=",
               "traveltime_alpha=", "parameters="]


def _g(x):
    """`out << double` with the default precision (6) and floatfield: printf %g."""
    return "%g" % x


class _Tok:
    """Whitespace-separated tokens; '{' and '}' are tokens of their own wherever they stand (the readers take brackets
    with `in >> char`, so `4{ 1 2}` and `{1.0` are legal)."""

    def __init__(self, path):
        with open(path) as f:
            self.t = re.findall(r"[{}]|[^\s{}]+", f.read())
        self.i, self.path = 0, path

    def next(self):
        if self.i >= len(self.t):
            raise ValueError(f"{self.path}: unexpected end of file")
        self.i += 1
        return self.t[self.i - 1]

    def peek(self):
        return self.t[self.i] if self.i < len(self.t) else None

    def expect(self, word):
        got = self.next()
        if got != word:
            raise ValueError(f"{self.path}: stuck at {got!r} instead of {word!r}")

    def int(self):
        return int(self.next())

    def float(self):
        return float(self.next())

    def count(self, keyword):
        self.expect(keyword)
        return self.int()


def read_master(path):
    """Network::readmaster: a fixed-order `key= value` list. Returns {key without '=': value}."""
    t = _Tok(path)
    t.expect("#input_files")
    out = {}
    for k in MASTER_KEYS:
        t.expect(k)
        if not k.startswith("#"):
            out[k[:-1]] = t.next()
    for k in ("starttime", "stoptime", "traveltime_alpha"):
        out[k] = float(out[k])
    out["calc_paths"] = int(out["calc_paths"])
    return out


def read_parameters(path):
    """parameters.dat: `#section` lines and `key= value` pairs in the order of Parameters::read_parameters; every pair is
    returned (numbers converted), the engine uses the kernel-relevant ones."""
    t = _Tok(path)
    out = {}
    while t.peek() is not None:
        k = t.next()
        if k.startswith("#"):
            continue
        if not k.endswith("="):
            raise ValueError(f"{path}: stuck at {k!r} instead of a `key=`")
        v = t.next()
        try:
            out[k[:-1]] = int(v)
        except ValueError:
            try:
                out[k[:-1]] = float(v)
            except ValueError:
                out[k[:-1]] = v
    return out


def read_network(path):
    """net.dat → (servers, nodes, sdfuncs, links) in NetSpec form."""
    t = _Tok(path)
    servers = []
    for _ in range(t.count("servers:")):
        t.expect("{")
        servers.append((t.int(), t.int(), t.float(), t.float(), t.float()))
        t.expect("}")
    nodes = []
    for _ in range(t.count("nodes:")):
        t.expect("{")
        nid, typ, x, y = t.int(), t.int(), t.float(), t.float()
        if typ not in (1, 2, 3):
            raise NotImplementedError(f"{path}: node {nid} has type {typ} (boundary nodes belong to the MIME/VISSIM hybrid)")
        sid = t.int() if typ == 2 else None
        t.expect("}")
        nodes.append((nid, typ, x, y, sid))
    sdfuncs = []
    for _ in range(t.count("sdfuncs:")):
        t.expect("{")
        sid, typ = t.int(), t.int()
        vals = [t.float() for _ in range(6 if typ in (1, 2) else 4)]
        t.expect("}")
        # (type 2 is a DynamitSdfunc as well, B/network.cpp:690-693)
        sdfuncs.append((sid, typ) + tuple(vals))
    links = []
    for _ in range(t.count("links:")):
        t.expect("{")
        rec = [t.int() for _ in range(6)]
        while t.peek() != "}":
            t.next()  # optional link name
        t.expect("}")
        links.append(tuple(rec))
    return servers, nodes, sdfuncs, links


def read_turnings(path):
    t = _Tok(path)
    turnings = []
    for _ in range(t.count("turnings:")):
        t.expect("{")
        turnings.append(tuple(t.int() for _ in range(6)))
        t.expect("}")
    if t.peek() == "giveways:" and t.count("giveways:") > 0:
        raise NotImplementedError(f"{path}: give-ways are not built (use_giveway is off by default)")
    return turnings


def read_demand(path):
    """demand.dat → ([(oid, did, rate·scale)] in file order (the engine sorts them like od_less_than),
    [(loadtime, scale, [(oid, did, rate)])] = the slices, i.e. the MatrixActions of Network::readrates, B/network.cpp:5544-5606)."""
    t = _Tok(path)
    n = t.count("od_pairs:")
    t.expect("scale:")
    scale = t.float()
    ods = []
    for _ in range(n):
        t.expect("{")
        o, d, rate = t.int(), t.int(), t.float()
        t.expect("}")
        ods.append((o, d, rate * scale))
    slices = []
    if t.peek() == "slices:":
        for _ in range(t.count("slices:")):
            m = t.count("od_pairs:")
            t.expect("scale:")
            sc = t.float()
            t.expect("loadtime:")
            lt = t.float()
            rates = []
            for _ in range(m):
                t.expect("{")
                rates.append((t.int(), t.int(), t.int()))  # readrate reads the rate as an int (B/network.cpp:5518-5522)
                t.expect("}")
            slices.append((lt, sc, rates))
    return ods, slices


def read_routes(path):
    """routes.dat → [(rid, oid, did, [link ids])], file order. A missing or empty file returns [] (the reference then
    forces calc_paths, B/network.cpp:7328-7332)."""
    if not os.path.exists(path):
        return []
    t = _Tok(path)
    if t.peek() != "routes:":
        return []
    routes = []
    for _ in range(t.count("routes:")):
        t.expect("{")
        rid, o, d, n = t.int(), t.int(), t.int(), t.int()
        t.expect("{")
        links = [t.int() for _ in range(n)]
        t.expect("}")
        t.expect("}")
        routes.append((rid, o, d, links))
    return routes


def read_linktimes(path):
    """histtimes.dat → (n_periods, periodlength, {lid: [P times]} or None for `links: 0` = free-flow times)."""
    t = _Tok(path)
    n = t.count("links:")
    P = t.count("periods:")
    t.expect("periodlength:")
    pl = t.float()
    if n == 0:
        return P, pl, None
    h = {}
    for _ in range(n):
        t.expect("{")
        lid = t.int()
        h[lid] = [t.float() for _ in range(P)]
        t.expect("}")
    return P, pl, h


def read_vtypes(path):
    t = _Tok(path)
    vt = []
    for _ in range(t.count("vtypes:")):
        t.expect("{")
        vt.append((t.int(), t.next(), t.float(), t.float()))
        t.expect("}")
    return vt


def read_serverrates(path):
    """serverrates.dat → [(server id, time, mu, sd)] in file order (Network::readserverrates, B/network.cpp:5608-5665)."""
    t = _Tok(path)
    out = []
    for _ in range(t.count("rates:")):
        t.expect("{")
        out.append((t.int(), t.float(), t.float(), t.float()))
        t.expect("}")
    return out


def read_signals(path):
    """signal.dat → [(control id, [(plan id, start, stop, offset, cycletime, [(stage id, start, duration, [turning ids])])])]
    (Network::readsignalcontrols / readsignalplan / readstage, B/network.cpp:5144-5295)."""
    t = _Tok(path)
    out = []
    for _ in range(t.count("controls:")):
        t.expect("{")
        cid, n_plans = t.int(), t.int()
        plans = []
        for _ in range(n_plans):
            t.expect("{")
            pid, start, stop, offset, cycle, n_stages = t.int(), t.float(), t.float(), t.float(), t.float(), t.int()
            stages = []
            for _ in range(n_stages):
                t.expect("{")
                sid, sstart, dur, n_turns = t.int(), t.float(), t.float(), t.int()
                t.expect("{")
                turns = [t.int() for _ in range(n_turns)]
                t.expect("}")
                t.expect("}")
                stages.append((sid, sstart, dur, turns))
            t.expect("}")
            plans.append((pid, start, stop, offset, cycle, stages))
        t.expect("}")
        out.append((cid, plans))
    return out


def read_incidents(path):
    """incident file (Network::readincidentfile, B/network.cpp:6337-6400, 6468-6480) → (its own sdfuncs, incidents as
    (link id, sdfunc id, penalty, start, stop, info_start, info_stop, blocked)). The parameters of the information
    broadcast that follow are not read: incidents with broadcast are refused by the flattening layer."""
    t = _Tok(path)
    sdfuncs = []
    for _ in range(t.count("sdfuncs:")):
        t.expect("{")
        sid, typ = t.int(), t.int()
        vals = [t.float() for _ in range(6 if typ in (1, 2) else 4)]
        t.expect("}")
        sdfuncs.append((sid, typ) + tuple(vals))
    inc = []
    for _ in range(t.count("incidents:")):
        t.expect("{")
        lid, sid = t.int(), t.int()
        penalty, start, stop, info_start, info_stop = (t.float() for _ in range(5))
        blocked = t.int()
        t.expect("}")
        inc.append((lid, sid, penalty, start, stop, info_start, info_stop, blocked))
    return sdfuncs, inc


def _require_zero(path, keyword, what):
    t = _Tok(path)
    while t.peek() is not None:
        if t.next() == keyword:
            if t.int() != 0:
                raise NotImplementedError(f"{path}: {what} are outside the two hot paths and not built")
            return
    raise ValueError(f"{path}: keyword {keyword!r} not found")


def read_scenario(masterfile, seed):
    """Network::readmaster + the readers of Network::executemaster → (NetSpec, master dict). `seed` is the run's
    randseed (the second command-line argument of mezzo_s; the engine needs an explicit one)."""
    wd = os.path.dirname(os.path.abspath(masterfile))
    m = read_master(masterfile)
    p = lambda k: os.path.join(wd, m[k])  # noqa: E731
    params = read_parameters(p("parameters"))
    vtypes = read_vtypes(p("vehicletypes"))  # (several classes: flatten.trip_lengths draws one per trip like ODaction::execute)
    if not vtypes:
        raise ValueError("vehiclemix.dat holds no vehicle class")
    servers, nodes, sdfuncs, links = read_network(p("network"))
    _require_zero(p("virtuallinks"), "virtuallinks:", "virtual links")
    serverrates = read_serverrates(p("serverrates"))
    turnings = read_turnings(p("turnings"))
    P, pl, hist = read_linktimes(p("histtimes"))
    ods, slices = read_demand(p("demand"))
    routes = read_routes(p("routes"))
    signals = read_signals(p("signals"))
    incident_sdfuncs, incidents = read_incidents(p("incident"))
    spec = dict(servers=servers, nodes=nodes, sdfuncs=sdfuncs, links=links, turnings=turnings, ods=ods, slices=slices, routes=routes,
                incident_sdfuncs=incident_sdfuncs, incidents=incidents,
                vtypes=vtypes, n_periods=P, periodlength=pl, histtimes=hist, params=params, serverrates=serverrates, signals=signals,
                starttime=m["starttime"], stoptime=m["stoptime"], randseed=int(seed))
    return spec, m


# ------------------------------------------------------------------------------------------------------- writers
def write_routes(path, routes):
    """Network::writepathfile: `routes` as returned by routes.shortest_paths_all (multimap<odval,Route*> order)."""
    with open(path, "w") as f:
        f.write("routes:\t%d\n" % len(routes))
        for rid, o, d, links in routes:
            f.write("{ %d %d %d %d{%s} }\n" % (rid, o, d, len(links), "".join(" %d" % l for l in links)))


def write_linktimes(path, link_ids, avg_final, hist, alpha, periodlength):
    """Network::writelinktimes: alpha·avg + (1−alpha)·hist per link (ascending id) and period; `avg_final` is
    Simulation.linktimes(final=True)."""
    L, P = avg_final.shape
    new = alpha * avg_final + (1 - alpha) * hist
    with open(path, "w") as f:
        f.write("links:\t%d\nperiods:\t%d\nperiodlength:\t%s\n" % (L, P, _g(periodlength)))
        for lid, row in zip(link_ids, new):
            f.write("{\t%d%s\t}\n" % (lid, "".join("\t" + _g(v) for v in row)))


OUTPUT_FIELDS = ["origin_id", "dest_id", "veh_id", "start_time", "end_time", "travel_time", "mileage", "route_id",
                 "switched_route"]


def write_output(path, od_rows):
    """Network::writeoutput: the field names, then every ODpair::grid row (OD pairs in od_less_than order, rows in
    arrival order) — `od_rows` is Simulation.od_rows()."""
    with open(path, "w") as f:
        f.write("".join(n + "\t" for n in OUTPUT_FIELDS) + "\n")
        for row in od_rows:
            f.write("".join(_g(v) + "\t" for v in row) + "\n")


def write_summary(path, od_list, od_rows):
    """Network::writesummary: one line per OD pair that kept routes — (origin, destination, nr of routes, arrived,
    Σ travel time, Σ mileage); the sums add in arrival order like Grid::sum. od_list = [(oid, did, n_routes)] sorted."""
    by = {}
    for row in od_rows:
        by.setdefault((int(row[0]), int(row[1])), []).append(row)
    with open(path, "w") as f:
        f.write("Origin\tDestination\tNrRoutes\tNrArrived\tTotalTravelTime(s)\tTotalMileage(m)\n")
        for o, d, nr in od_list:
            rows = by.get((o, d), [])
            tt = mil = 0.0
            for r in rows:
                tt += r[5]
                mil += r[6]
            f.write("%d\t%d\t%d\t%d\t%s\t%s\n" % (o, d, nr, len(rows), _g(tt), _g(mil)))


# ------------------------------------------------------------------------------------------------------- driver
def run_masterfile(masterfile, seed, device=0, replication=0):
    """Network::executemaster + Network::step + the link-time / OD part of Network::writeall for a scenario directory,
    with both hot paths on the GPU: route generation (when calc_paths=1 or routes.dat is empty) through mz_sssp_routes,
    the simulation through mz_sim_*. routes.dat is written back like the reference does. Returns (Simulation stats, spec)."""
    from . import Simulation, flatten, routes as mzroutes
    wd = os.path.dirname(os.path.abspath(masterfile))
    spec, m = read_scenario(masterfile, seed)
    if m["calc_paths"] or not spec["routes"]:
        fn = mzroutes.gpu_route_fn(spec, device=device)
        try:
            spec["routes"] = mzroutes.shortest_paths_all(spec, fn)
        finally:
            fn.graph.close()
    else:
        # the routemap is a multimap keyed by (origin, destination): writepathfile streams it in that order
        spec["routes"] = [r for _, r in sorted(enumerate(spec["routes"]), key=lambda x: (x[1][1], x[1][2], x[0]))]
    flat = flatten.FlatNet(spec)
    # "write back the routes" (B/network.cpp:7344-7351): what add_od_routes left in the routemap, i.e. without the routes
    # delete_spurious_routes threw out when delete_bad_routes=1
    spec["routes"] = [r for r, k in zip(spec["routes"], flat.route_kept) if k]
    write_routes(os.path.join(wd, m["routes"]), spec["routes"])
    sim = Simulation(flat, device=device)
    try:
        sim.step()
        stats = sim.stats()
        rep = (".%d" % replication) if replication > 0 else ""
        lt = os.path.join(wd, m["linktimes"])
        avg = sim.linktimes(final=True)
        write_linktimes(lt + rep, flat.link_ids, avg, flat.link_hist, m["traveltime_alpha"], spec["periodlength"])
        write_linktimes(lt + ".clean" + rep, flat.link_ids, avg, flat.link_hist, 1.0, spec["periodlength"])
        rows = sim.od_rows()
        nr = {}
        for r in spec["routes"]:
            nr[(r[1], r[2])] = nr.get((r[1], r[2]), 0) + 1
        od_list = [(o, d, nr[(o, d)]) for o, d, _ in flat.od_sorted if nr.get((o, d), 0) > 0]
        write_summary(os.path.join(wd, m["summary"]) + rep, od_list, rows)
        write_output(os.path.join(wd, m["output"]) + rep, rows)
    finally:
        sim.close()
    return stats, spec
