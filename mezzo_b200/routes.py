"""Host mirror of the route-generation DRIVER around the SSSP kernels: Network::init_shortest_path +
Network::shortest_paths_all + exists_same_route (B/network.cpp:6576-6630, 6667-6774, 458-470;
B/ = /root/reference/branches/busmezzo/mezzo_lib/src/).

The reference calls labelCorrecting once per (rerun, origin, outgoing link) and walks the destinations after each tree;
here ALL trees of a run are handed to the device in one batch (mz_sssp_routes: trees + reachable + path walk + root
prepend on the GPU) and the route bookkeeping — route numbering, de-duplication per OD pair, multimap order — is replayed
on the host in the reference's loop order, so route ids and route sets come out exactly as `routes.dat` rewritten by the
reference (Network::writepathfile, B/network.cpp:6505-6517)."""
import numpy as np

SP_UNSET = -2


def graph_arrays(spec):
    """Graph<double,LinkTimeInfo>(max node id+1, max link id+1) as Network::init_shortest_path builds it: ids are array
    slots, links added in ascending id order, turnings with sid < 0 are turning prohibitors. Returns
    (n_node_slots, n_link_slots, up, dn, prohibitors, T[slots,P], P, periodlength)."""
    links = sorted(spec["links"], key=lambda l: l[0])
    n_node_slots = max(n[0] for n in spec["nodes"]) + 1
    n_link_slots = links[-1][0] + 1
    up = np.full(n_link_slots, SP_UNSET, np.int32)
    dn = np.full(n_link_slots, SP_UNSET, np.int32)
    for l in links:
        up[l[0]], dn[l[0]] = l[1], l[2]
    prohibitors = [(t[3], t[4]) for t in spec["turnings"] if t[2] < 0]
    P = int(spec["n_periods"])
    T = np.full((n_link_slots, P), np.nan)
    if spec.get("histtimes") is None:
        # "links: 0" in histtimes.dat: every period = the link's free-flow time (B/network.cpp:6254-6273)
        sd = {s[0]: s for s in spec["sdfuncs"]}
        for l in links:
            T[l[0], :] = max(1.0, l[3] / sd[l[5]][2])
    else:
        for lid, row in spec["histtimes"].items():
            T[int(lid), :len(row)] = row
    return n_node_slots, n_link_slots, up, dn, prohibitors, T, P, float(spec["periodlength"])


def shortest_paths_all(spec, route_fn):
    """Returns the route list [(rid, oid, did, [link ids])] the reference holds in `routemap` after shortest_paths_all:
    the routes of spec["routes"] plus every new shortest path, ordered like the multimap<odval,Route*> (by (origin,
    destination), insertion order within an OD pair).

    route_fn(roots, entries, dest_off, dests) -> (path_off, path_links) computes, for each tree, the path to each listed
    destination NODE (empty if unreachable), root link prepended — mezzo_b200.Graph.sssp_routes on a GPU."""
    interval = float(spec["params"].get("update_interval_routes", 900.0))
    nr_reruns = int(spec["stoptime"] / interval) - 1  # "except for last period"
    odpairs = sorted(spec["ods"], key=lambda x: (x[0], x[1]))  # sort(odpairs.begin(), odpairs.end(), od_less_than)
    out_links = {}
    for l in sorted(spec["links"], key=lambda l: l[0]):
        out_links.setdefault(l[1], []).insert(0, l[0])  # Origin::register_links: outgoing.insert(begin()) ⇒ descending id
    # routemap: OD -> routes in insertion order
    routemap = {}
    for r in spec["routes"]:
        routemap.setdefault((r[1], r[2]), []).append((r[0], list(r[3])))
    routenr = len(spec["routes"])
    # the trees of the whole run, in the reference's loop order
    roots, entries, dest_off, dests, jobs = [], [], [0], [], []
    for i in range(nr_reruns):
        entry = i * interval
        k = 0
        while k < len(odpairs):
            origin = odpairs[k][0]
            group = []
            while k < len(odpairs) and odpairs[k][0] == origin:
                # at this point no OD pair has routes attached yet (add_od_routes runs afterwards), so nr_routes < 1 holds
                # and every destination of the origin is searched (B/network.cpp:6690-6694)
                group.append(odpairs[k][1])
                k += 1
            for root in out_links.get(origin, []):
                roots.append(root)
                entries.append(entry)
                dests.extend(group)
                dest_off.append(len(dests))
                jobs.append((origin, root, group))
    if not roots:
        return [(rid, o, d, links) for (o, d), rs in sorted(routemap.items()) for rid, links in rs]
    path_off, path_links = route_fn(np.asarray(roots, np.int32), np.asarray(entries, np.float64),
                                    np.asarray(dest_off, np.int64), np.asarray(dests, np.int32))
    pair = 0
    for origin, root, group in jobs:
        for d in group:
            path = [int(x) for x in path_links[path_off[pair]:path_off[pair + 1]]]
            pair += 1
            if not path:
                continue  # not reachable from this root link (or circular: "PROBLEM OBTAINING links")
            routenr += 1  # consumed even when the route turns out to exist already
            have = routemap.setdefault((origin, d), [])
            if not any(links == path for _, links in have):  # exists_same_route: same OD, same link sequence
                have.append((routenr, path))
    return [(rid, o, d, links) for (o, d), rs in sorted(routemap.items()) for rid, links in rs]


def gpu_route_fn(spec, device=0, mode=None):
    """route_fn backed by the engine: builds the graph + link-time table on `device` once."""
    from . import MZ_SSSP_AUTO, Graph
    n_nodes, n_links, up, dn, pro, T, P, pl = graph_arrays(spec)
    g = Graph.from_arrays(n_nodes, n_links, up, dn, prohibitors=pro, device=device)
    g.set_linktimes(T, P, pl)
    m = MZ_SSSP_AUTO if mode is None else mode

    def fn(roots, entries, dest_off, dests):
        return g.sssp_routes(roots, entries, dest_off, dests, mode=m)

    fn.graph = g
    return fn
