// Backend-independent host logic of the link-update engine: validation, flattening of Network::init's first events
// into per-node action tables, state reset, the super-step driver loop and the result getters.
// B is the memory/launch backend: CudaBackend in sim.cu (the product) or HostBackend in tests/emu/sim_emu.cpp
// (test-only emulation). Reference citations: see sim_core.h.
#pragma once
#include <algorithm>
#include <chrono>
#include <cstddef>
#include <cmath>
#include <cstdio>
#include <condition_variable>
#include <cstdlib>
#include <deque>
#include <map>
#include <mutex>
#include <set>
#include <string>
#include <thread>
#include <vector>

#include "errors.h"
#include "sim_core.h"

namespace mzsim {

// std::vector whose resize() leaves trivially constructible elements untouched: the large tables of the build are
// filled — and their pages first touched — by the worker threads, not zeroed by the calling thread beforehand.
template <class T>
struct NoInit : std::allocator<T> {
    template <class U> struct rebind { using other = NoInit<U>; };
    template <class U, class... A>
    void construct(U *p, A &&...a)
    {
        if constexpr (sizeof...(A) == 0) ::new ((void *)p) U;
        else ::new ((void *)p) U(std::forward<A>(a)...);
    }
};
template <class T> using RawVector = std::vector<T, NoInit<T>>;

// Partition of the nodes over ranks (GPUs). node_rank == nullptr / n_ranks == 1: everything on one rank.
struct SimPart {
    int n_ranks = 1, rank = 0;
    const int32_t *node_rank = nullptr;  // [n_nodes]
};

template <class B>
struct SimHost {
    int device = 0;
    SimDev dev{};
    std::vector<void *> allocs;
    // initial images of the mutable state (uploaded by sim_reset_state)
    RawVector<ActDyn> h_adyn;
    std::vector<double> h_key_t, h_key_tp, h_link_hist;
    std::vector<unsigned char> h_turn_active0;
    long long n_sig_events = 0;
    int n_sig_controls = 0;
    RawVector<int> h_act_node;
    std::vector<int> h_route_off, h_route_links, h_trip_route, h_otrip_off;
    int n_trips = 0, n_links = 0, n_periods = 0, n_nodes = 0, n_acts = 0, n_origins = 0, n_servers_total = 0, n_odpairs = 0;
    int n_ranks = 1, rank = 0;
    long long total_log = 0, slab_entries = 0;
    long long randseed = 0;
    long long n_init_events = 0;
    double phantom_t = 0.0, phantom_tp = 0.0;  // MzNet::phantom_final_t / _tp
    std::vector<long long> h_trip_log_off;
    std::vector<unsigned char> h_node_home;
    std::vector<int> h_node_rank;
    size_t off_sync = 0;  // arena offset of the cross-rank window barrier words
    void *arena = nullptr;  // this rank's arena (shared mutable state)
    size_t arena_bytes = 0;
    typename B::Ctx ctx;  // streams / events of the backend
    mz_sim_stats stats{};
    bool ran = false;
    int blocks_override = 0;
    int kernel_variant = 0;  // MZ_SIM_OPT_KERNEL: 0 automatic, 1 warp per node, 2 thread per node
    unsigned long long prof_ev[56] = {0};  // MZ_SIM_PROFILE builds: per event class {cycles, count, max cycles}
    bool peers_attached = false;
    // regions: this rank's nodes split into contiguous pieces, one per thread block of the node-visit kernel (sim_set_regions)
    std::vector<int> h_nb_off, h_nb, h_link_in, h_link_out, h_my_nodes;
    std::vector<double> h_node_weight;
    int n_regions = 0, max_region = 0, region_cap = 0;
    std::vector<int> h_region_pre;  // sim_build's worker thread: region of every node for `pre_regions` regions
    int pre_regions = 0;
    int *d_my_nodes = nullptr, *d_region_off = nullptr, *d_node_nb = nullptr;
    int *d_nb_slot = nullptr, *d_mb_off = nullptr, *d_poll_off = nullptr, *d_poll_pos = nullptr;
    size_t mailbox_words = 0;
    // pristine images of the mutable state that is not a plain byte pattern, kept on the device: Network::reset
    // (mz_sim_reset) is then device-to-device copies and fills, not a re-upload
    ActDyn *d_adyn0 = nullptr;
    double *d_key_t0 = nullptr;
    KeySlot *d_kslot0 = nullptr;
    unsigned char *d_turn_active0 = nullptr;
    bool init_images = false;
    unsigned char *d_node_cls = nullptr, *d_link_cls = nullptr;
    // transit overlay hooks (row f3): stop links, their ACT_STOP actions, the departures the host model has booked
    struct Dep {
        double t;
        long long seq;
        int veh, slink;
    };
    int n_stops = 0, n_slinks = 0;
    std::vector<int> h_trip_stop_off, h_stop_link, h_slink_link, h_slink_act, h_bus_slink;  // h_bus_slink: by first stop slot, -1 = not at a stop
    std::vector<Dep> pending;
    long long dep_seq = 0, visits_read = 0;
    double ran_until = 0.0, horizon = 0.0, accum_ms = 0.0;
    int *d_dep_off = nullptr, *d_dep_veh = nullptr;
    double *d_dep_t = nullptr;
};

template <class B, class T, class Al>
int dev_upload(SimHost<B> *s, const T *&dst, const std::vector<T, Al> &src)
{
    void *p = nullptr;
    if (int rc = B::alloc(s->ctx, &p, std::max<size_t>(1, src.size()) * sizeof(T))) return rc;
    s->allocs.push_back(p);
    if (!src.empty())
        if (int rc = B::h2d(s->ctx, p, src.data(), src.size() * sizeof(T))) return rc;
    dst = (const T *)p;
    return MZ_OK;
}
template <class B, class T>
int dev_alloc(SimHost<B> *s, T *&dst, size_t n)
{
    void *p = nullptr;
    if (int rc = B::alloc(s->ctx, &p, std::max<size_t>(1, n) * sizeof(T))) return rc;
    s->allocs.push_back(p);
    dst = (T *)p;
    return MZ_OK;
}
// Device room for a table one of the build's worker threads uploads itself as soon as it has filled it (B::h2d_worker);
// allocation stays on the calling thread (the allocator and s->allocs are not shared between threads).
template <class B, class T>
int dev_reserve(SimHost<B> *s, const T *&dst, size_t n)
{
    void *p = nullptr;
    if (int rc = B::alloc(s->ctx, &p, std::max<size_t>(1, n) * sizeof(T))) return rc;
    s->allocs.push_back(p);
    dst = (const T *)p;
    return MZ_OK;
}
template <class T>
std::vector<T> vec(const T *p, size_t n)
{
    return std::vector<T>(p, p + n);
}

inline double host_sd_speed0(const MzNet *n, int f)
{
    // Sdfunc::speed(0.0): 0 <= romin always (reader asserts romin >= 0) ⇒ vfree
    return 0.0 <= n->sd_romin[f] ? n->sd_vfree[f] : n->sd_vmin[f];
}

inline int sim_validate(const MzNet *n)
{
    MZ_REQUIRE(n, MZ_EINVAL, "null network");
    MZ_REQUIRE(n->n_nodes > 0 && n->n_links > 0 && n->n_sdfuncs > 0 && n->n_periods > 0, MZ_EINVAL,
               "network needs nodes, links, sdfuncs and link-time periods");
    MZ_REQUIRE(n->randseed != 0, MZ_EINVAL, "randseed must be != 0 (0 means clock seeding in the reference)");
    MZ_REQUIRE(n->use_giveway == 0, MZ_ELIMIT, "use_giveway is not supported (SURVEY.md: off by default)");
    MZ_REQUIRE(n->standard_veh_length > 0 && n->periodlength > 0 && n->runtime > 0, MZ_EINVAL, "bad parameters");
    for (int i = 0; i < n->n_links; ++i) {
        MZ_REQUIRE(n->link_length[i] > 0 && n->link_lanes[i] > 0, MZ_EINVAL, "link %d: length and lanes must be > 0", i);
        MZ_REQUIRE(n->link_sdfunc[i] >= 0 && n->link_sdfunc[i] < n->n_sdfuncs, MZ_EINVAL, "link %d: bad sdfunc", i);
        MZ_REQUIRE(n->link_in_node[i] >= 0 && n->link_in_node[i] < n->n_nodes && n->link_out_node[i] >= 0 &&
                       n->link_out_node[i] < n->n_nodes, MZ_EINVAL, "link %d: bad end node", i);
    }
    for (int i = 0; i < n->n_servers; ++i)
        // type 3 (LogNormalDelayServer) draws its delay from the GLOBAL theRandomizers[0] (B/server.cpp:132-137): its stream
        // follows the global event order, which no partitioned execution can reproduce
        MZ_REQUIRE(n->srv_type[i] >= 0 && n->srv_type[i] <= 4 && n->srv_type[i] != 3, MZ_ELIMIT,
                   "server %d: type %d not supported (0, 1, 2 and 4 are)", i, n->srv_type[i]);
    for (int k = 0; k < n->n_turnings; ++k) {
        MZ_REQUIRE(n->turn_in[k] >= 0 && n->turn_in[k] < n->n_links && n->turn_out[k] >= 0 && n->turn_out[k] < n->n_links,
                   MZ_EINVAL, "turning %d: bad link", k);
        MZ_REQUIRE(n->turn_server[k] >= 0 && n->turn_server[k] < n->n_servers, MZ_EINVAL, "turning %d: bad server", k);
        MZ_REQUIRE(n->turn_in[k] != n->turn_out[k], MZ_ELIMIT, "turning %d turns a link into itself", k);
        MZ_REQUIRE(n->link_out_node[n->turn_in[k]] == n->turn_node[k] && n->link_in_node[n->turn_out[k]] == n->turn_node[k],
                   MZ_EINVAL, "turning %d: its links do not meet at its node", k);
    }
    for (int i = 0; i < n->n_incidents; ++i) {
        MZ_REQUIRE(n->inc_link && n->inc_sdfunc && n->inc_start && n->inc_stop, MZ_EINVAL, "incident arrays missing");
        MZ_REQUIRE(n->inc_link[i] >= 0 && n->inc_link[i] < n->n_links, MZ_EINVAL, "incident %d: bad link", i);
        MZ_REQUIRE(n->inc_sdfunc[i] >= 0 && n->inc_sdfunc[i] < n->n_sdfuncs, MZ_EINVAL, "incident %d: bad sdfunc", i);
        MZ_REQUIRE(n->inc_start[i] > 0.0 && n->inc_stop[i] > n->inc_start[i], MZ_EINVAL, "incident %d: needs 0 < start < stop", i);
    }
    {
        std::vector<int> seen((size_t)n->n_links, -1);  // last route that used the link
        for (int r = 0; r < n->n_routes; ++r) {
            MZ_REQUIRE(n->route_off[r] < n->route_off[r + 1], MZ_EINVAL, "route %d is empty", r);
            for (int i = n->route_off[r]; i < n->route_off[r + 1]; ++i) {
                const int l = n->route_links[i];
                MZ_REQUIRE(l >= 0 && l < n->n_links, MZ_EINVAL, "route %d: bad link", r);
                // Route::nextlink searches the FIRST occurrence of the current link (B/route.cpp:139-149)
                MZ_REQUIRE(seen[(size_t)l] != r, MZ_ELIMIT, "route %d visits link %d twice", r, l);
                seen[(size_t)l] = r;
            }
        }
    }
    if (n->trip_stop_off) {
        MZ_REQUIRE(n->stop_link && n->stop_pos, MZ_EINVAL, "trip_stop_off without stop_link / stop_pos");
        MZ_REQUIRE(n->trip_stop_off[0] == 0, MZ_EINVAL, "trip_stop_off[0] must be 0");
        for (int v = 0; v < n->n_trips; ++v) {
            const int s0 = n->trip_stop_off[v], s1 = n->trip_stop_off[v + 1];
            MZ_REQUIRE(s0 <= s1, MZ_EINVAL, "trip %d: stop offsets decrease", v);
            const int r = n->trip_route[v];
            int ri = n->route_off[r];
            for (int k = s0; k < s1; ++k) {
                const int l = n->stop_link[k];
                MZ_REQUIRE(l >= 0 && l < n->n_links, MZ_EINVAL, "trip %d, stop %d: bad link", v, k - s0);
                MZ_REQUIRE(n->stop_pos[k] >= 0.0 && n->stop_pos[k] <= n->link_length[l], MZ_EINVAL,
                           "trip %d, stop %d: position outside its link", v, k - s0);
                if (k > s0 && n->stop_link[k - 1] == l) {
                    MZ_REQUIRE(n->stop_pos[k] >= n->stop_pos[k - 1], MZ_EINVAL, "trip %d: stops on link %d not in driving order", v, l);
                    continue;
                }
                while (ri < n->route_off[r + 1] && n->route_links[ri] != l) ++ri;
                MZ_REQUIRE(ri < n->route_off[r + 1], MZ_EINVAL,
                           "trip %d, stop %d: link %d does not follow on the trip's route (stops must be in route order)",
                           v, k - s0, l);
            }
        }
    }
    // the trip table (10 M rows at the target size) is scanned by four threads for its first offending row; the message
    // for that row comes from the sequential checks below
    int v_first = 0;
    if (n->n_trips > (1 << 16)) {
        const int T = 4;
        int first_bad[T];
        std::vector<std::thread> th;
        for (int t = 0; t < T; ++t)
            th.emplace_back([&, t] {
                const int v0 = (int)((long long)n->n_trips * t / T), v1 = (int)((long long)n->n_trips * (t + 1) / T);
                first_bad[t] = n->n_trips;
                for (int v = v0; v < v1; ++v) {
                    bool ok = n->trip_route[v] >= 0 && n->trip_route[v] < n->n_routes && n->trip_origin[v] >= 0 &&
                              n->trip_origin[v] < n->n_origins && (v == 0 || n->trip_time[v] >= n->trip_time[v - 1]) &&
                              n->trip_od[v] >= 0 && n->trip_od[v] < n->n_odpairs && n->trip_prev_time[v] <= n->trip_time[v];
                    ok = ok && n->link_in_node[n->route_links[n->route_off[n->trip_route[v]]]] == n->origin_node[n->trip_origin[v]];
                    if (!ok) {
                        first_bad[t] = v;
                        break;
                    }
                }
            });
        for (std::thread &t : th) t.join();
        v_first = n->n_trips;
        for (int t = T - 1; t >= 0; --t)
            if (first_bad[t] < n->n_trips) v_first = first_bad[t];
    }
    for (int v = v_first; v < n->n_trips; ++v) {
        MZ_REQUIRE(n->trip_route[v] >= 0 && n->trip_route[v] < n->n_routes, MZ_EINVAL, "trip %d: bad route", v);
        MZ_REQUIRE(n->trip_origin[v] >= 0 && n->trip_origin[v] < n->n_origins, MZ_EINVAL, "trip %d: bad origin", v);
        MZ_REQUIRE(v == 0 || n->trip_time[v] >= n->trip_time[v - 1], MZ_EINVAL, "trips must be in time order");
        MZ_REQUIRE(n->trip_od[v] >= 0 && n->trip_od[v] < n->n_odpairs, MZ_EINVAL, "trip %d: bad OD pair", v);
        MZ_REQUIRE(n->trip_prev_time[v] <= n->trip_time[v], MZ_EINVAL, "trip %d: booked after it happens", v);
        const int first = n->route_links[n->route_off[n->trip_route[v]]];
        MZ_REQUIRE(n->link_in_node[first] == n->origin_node[n->trip_origin[v]], MZ_EINVAL,
                   "trip %d: route does not start at its origin", v);
    }
    return MZ_OK;
}

// Graph-growing recursive bisection of the node set `ids` into `parts` pieces of (nearly) equal weight — the heuristic
// METIS uses for its initial partitions (no METIS in the image): breadth-first order from a peripheral node of the induced
// subgraph, cut where the running weight reaches the left share, recurse. On road networks the pieces are contiguous
// patches with short boundaries. `member` / `seen` are scratch arrays of n_nodes ints (stamped, never cleared).
inline void region_bisect(const std::vector<int> &nb_off, const std::vector<int> &nb, const std::vector<double> &weight,
                          std::vector<int> &ids, int r0, int parts, std::vector<int> &region_of, std::vector<int> &member,
                          std::vector<int> &seen, int &stamp)
{
    if (parts <= 1 || ids.size() <= 1) {
        for (int v : ids) region_of[v] = r0;
        return;
    }
    const int me = ++stamp;
    for (int v : ids) member[v] = me;
    std::vector<int> order;
    order.reserve(ids.size());
    auto bfs = [&](int start) {
        const int mark = ++stamp;
        order.clear();
        size_t next_seed = 0;
        int seed = start;
        for (;;) {
            if (seen[seed] != mark) {
                seen[seed] = mark;
                size_t head = order.size();
                order.push_back(seed);
                while (head < order.size()) {
                    const int u = order[head++];
                    for (int i = nb_off[u]; i < nb_off[u + 1]; ++i) {
                        const int v = nb[i];
                        if (member[v] == me && seen[v] != mark) {
                            seen[v] = mark;
                            order.push_back(v);
                        }
                    }
                }
            }
            while (next_seed < ids.size() && seen[ids[next_seed]] == mark) ++next_seed;  // further components
            if (next_seed == ids.size()) break;
            seed = ids[next_seed];
        }
    };
    bfs(ids[0]);
    bfs(order.back());  // start again from a peripheral node
    const int left_parts = parts / 2;
    double total = 0.0;
    for (int v : order) total += weight[v];
    const double target = total * left_parts / parts;
    size_t cut = 0;
    double run = 0.0;
    while (cut < order.size() && run < target) run += weight[order[cut++]];
    cut = std::min(std::max(cut, (size_t)left_parts), order.size() - (size_t)(parts - left_parts));
    std::vector<int> left(order.begin(), order.begin() + (long)cut), right(order.begin() + (long)cut, order.end());
    std::vector<int>().swap(order);
    std::vector<int>().swap(ids);
    region_bisect(nb_off, nb, weight, left, r0, left_parts, region_of, member, seen, stamp);
    region_bisect(nb_off, nb, weight, right, r0 + left_parts, parts - left_parts, region_of, member, seen, stamp);
}

// node neighbours (nodes sharing a link) as CSR, each list ascending and without duplicates
inline void node_neighbours(const MzNet *n, std::vector<int> &nb_off, std::vector<int> &nb)
{
    const int N = n->n_nodes, L = n->n_links;
    nb_off.assign((size_t)N + 1, 0);
    nb.clear();
    // counting sort of the (node, other end) pairs, then sort + unique inside each node's short list
    std::vector<int> cnt(N + 1, 0);
    for (int l = 0; l < L; ++l) {
        const int a = n->link_in_node[l], b = n->link_out_node[l];
        if (a != b) {
            cnt[a + 1]++;
            cnt[b + 1]++;
        }
    }
    for (int i = 0; i < N; ++i) cnt[i + 1] += cnt[i];
    std::vector<int> all((size_t)cnt[N]), pos(cnt.begin(), cnt.end() - 1);
    for (int l = 0; l < L; ++l) {
        const int a = n->link_in_node[l], b = n->link_out_node[l];
        if (a != b) {
            all[(size_t)pos[a]++] = b;
            all[(size_t)pos[b]++] = a;
        }
    }
    nb.reserve(all.size());
    for (int i = 0; i < N; ++i) {
        std::sort(all.begin() + cnt[i], all.begin() + cnt[i + 1]);
        for (int k = cnt[i]; k < cnt[i + 1]; ++k)
            if (k == cnt[i] || all[(size_t)k] != all[(size_t)k - 1]) nb.push_back(all[(size_t)k]);
        nb_off[i + 1] = (int)nb.size();
    }
}

// Node partition for a multi-GPU engine (mz_partition_nodes): the same graph-growing recursive bisection that cuts a rank's
// nodes into regions, here over all nodes, balanced by the expected load of a node (every trip crosses the down node of
// each link of its route once). METIS-style initial partition; any other node_rank[] is as valid for mz_sim_create_part.
inline int partition_nodes(const MzNet *n, int n_ranks, int32_t *node_rank)
{
    MZ_REQUIRE(n && node_rank, MZ_EINVAL, "null argument");
    MZ_REQUIRE(n_ranks >= 1 && n_ranks <= MZ_MAXR, MZ_EINVAL, "n_ranks must be in [1, %d]", MZ_MAXR);
    if (int rc = sim_validate(n)) return rc;
    const int N = n->n_nodes;
    std::vector<int> nb_off, nb;
    node_neighbours(n, nb_off, nb);
    std::vector<double> weight((size_t)N, 1.0);
    {
        std::vector<int> route_trips(std::max(1, n->n_routes), 0);
        for (int v = 0; v < n->n_trips; ++v) route_trips[n->trip_route[v]]++;
        for (int r = 0; r < n->n_routes; ++r)
            if (route_trips[r])
                for (int i = n->route_off[r]; i < n->route_off[r + 1]; ++i)
                    weight[(size_t)n->link_out_node[n->route_links[i]]] += 0.02 * route_trips[r];
    }
    std::vector<int> ids((size_t)N), region_of((size_t)N, 0), member((size_t)N, 0), seen((size_t)N, 0);
    for (int i = 0; i < N; ++i) ids[(size_t)i] = i;
    int stamp = 0;
    region_bisect(nb_off, nb, weight, ids, 0, n_ranks, region_of, member, seen, stamp);
    for (int i = 0; i < N; ++i) node_rank[i] = region_of[(size_t)i];
    return MZ_OK;
}

// Stable grouping of the trips by a key (origin, OD pair): off[k] = first position of key k, list = trip ids in ascending
// order inside each key, slot[v] = position of trip v inside its key (optional). A counting sort whose two passes over the
// trip table are cut into four contiguous pieces, one thread each.
template <class V>
void group_trips(int n_trips, const int32_t *key, int n_keys, std::vector<int> &off, V &list, V *slot)
{
    const int T = n_trips > (1 << 18) ? 4 : 1;
    std::vector<std::vector<int>> cnt((size_t)T, std::vector<int>((size_t)std::max(1, n_keys), 0));
    auto piece = [&](int t, int &v0, int &v1) {
        v0 = (int)((long long)n_trips * t / T);
        v1 = (int)((long long)n_trips * (t + 1) / T);
    };
    auto run = [&](auto body) {
        std::vector<std::thread> th;
        for (int t = 1; t < T; ++t) th.emplace_back(body, t);
        body(0);
        for (std::thread &x : th) x.join();
    };
    run([&](int t) {
        int v0, v1;
        piece(t, v0, v1);
        std::vector<int> &c = cnt[(size_t)t];
        for (int v = v0; v < v1; ++v) c[key[v]]++;
    });
    off.assign((size_t)n_keys + 1, 0);
    for (int k = 0; k < n_keys; ++k) {
        int at = off[(size_t)k];
        for (int t = 0; t < T; ++t) {  // cnt[t][k] becomes the first position of piece t inside key k
            const int c = cnt[(size_t)t][(size_t)k];
            cnt[(size_t)t][(size_t)k] = at;
            at += c;
        }
        off[(size_t)k + 1] = at;
    }
    list.resize((size_t)n_trips);
    if (slot) slot->resize((size_t)n_trips);
    run([&](int t) {
        int v0, v1;
        piece(t, v0, v1);
        std::vector<int> &c = cnt[(size_t)t];
        for (int v = v0; v < v1; ++v) {
            const int k = key[v], at = c[k]++;
            list[(size_t)at] = v;
            if (slot) (*slot)[(size_t)v] = at - off[(size_t)k];
        }
    });
}

template <class B> int sim_reset_state(SimHost<B> *s);
template <class B> int sim_set_regions(SimHost<B> *s, int n_regions);
template <class B> int sim_destroy(SimHost<B> *s);
template <class B> int sim_min_key(SimHost<B> *s, Key *out);

template <class B>
int sim_build(SimHost<B> *s, const MzNet *n)
{
    SimDev &d = s->dev;
    const int L = n->n_links, N = n->n_nodes, P = n->n_periods;
    auto t_last = std::chrono::steady_clock::now();
    const bool timing = std::getenv("MZ_TIMING") != nullptr;
    auto tick = [&](const char *what) {
        if (!timing) return;
        auto now = std::chrono::steady_clock::now();
        std::fprintf(stderr, "[mz timing] %-28s %8.2f ms\n", what, std::chrono::duration<double, std::milli>(now - t_last).count());
        t_last = now;
    };
    s->n_links = L;
    s->n_nodes = N;
    s->n_periods = P;
    s->n_trips = n->n_trips;
    s->n_origins = n->n_origins;
    s->n_odpairs = n->n_odpairs;
    s->n_sig_controls = n->n_sig_controls;
    s->randseed = n->randseed;
    s->phantom_t = n->phantom_final_t;
    s->phantom_tp = n->phantom_final_tp;
    d.n_nodes = N;
    d.n_links = L;
    d.n_turnings = n->n_turnings;
    d.n_trips = n->n_trips;
    d.n_periods = P;
    d.max_events = 1 << 30;  // no cap by default (the cap only bounds the latency of one visit; measured best on B200)
    d.runtime = n->runtime;
    d.periodlength = n->periodlength;
    d.server_type = n->server_type;

    // ---- links: constants of the Link constructor (B/link.cpp:8-49)
    std::vector<int> maxcap(L), qoff(L), qcap(L);
    std::vector<double> freeflow(L);
    long long slab = 0;
    for (int i = 0; i < L; ++i) {
        maxcap[i] = (int)(n->link_length[i] * n->link_lanes[i] / n->standard_veh_length);  // int arithmetic
        double ff = n->link_length[i] / host_sd_speed0(n, n->link_sdfunc[i]);
        if (ff < 1.0) ff = 1.0;
        freeflow[i] = ff;
        qcap[i] = std::max(1, maxcap[i]);
        MZ_REQUIRE(slab + qcap[i] < (long long)INT32_MAX, MZ_ELIMIT, "queue slab exceeds 2^31 entries");
        qoff[i] = (int)slab;
        slab += qcap[i];
    }
    s->slab_entries = slab;
    s->h_link_hist = vec(n->link_hist, (size_t)L * P);

    for (int o = 0; o < n->n_origins; ++o)
        MZ_REQUIRE(n->origin_node[o] >= 0 && n->origin_node[o] < N, MZ_EINVAL, "origin %d: bad node", o);
    for (int k = 0; k < n->n_dests; ++k)
        MZ_REQUIRE(n->dest_node[k] >= 0 && n->dest_node[k] < N, MZ_EINVAL, "destination %d: bad node", k);

    tick("links");
    // Two pieces of the build are scatter passes over the trip table (10 M trips at the target size) that nothing before
    // the upload depends on: the trips per origin and the route / exit-log tables. They run on their own threads beside
    // the action build (joined before the upload; the joiner also covers every early return in between).
    struct Joiner {
        std::thread t;
        ~Joiner() { if (t.joinable()) t.join(); }
    };
    // The big per-trip tables (two thirds of the upload at the target size) go to the device from the threads that fill
    // them, while the action build below is still running; the first failure is kept for the check before the upload
    // section (error texts are thread-local).
    struct Early {
        std::mutex m;
        std::condition_variable cv;
        struct Job {
            const void *dst, *src;
            size_t bytes;
        };
        std::deque<Job> jobs;
        bool closed = false;
        int rc = MZ_OK;
        std::string msg;
    } early;
    // (called by the threads that fill the tables; the copies themselves are made by th_uploader, in the order queued)
    auto early_up = [&](const void *dst, const void *src, size_t bytes) {
        {
            std::lock_guard<std::mutex> g(early.m);
            early.jobs.push_back({dst, src, bytes});
        }
        early.cv.notify_one();
    };
    struct Uploader {
        Early &e;
        std::thread t;
        void finish()
        {
            if (!t.joinable()) return;
            {
                std::lock_guard<std::mutex> g(e.m);
                e.closed = true;
            }
            e.cv.notify_one();
            t.join();
        }
        ~Uploader() { finish(); }
    };
    Uploader th_uploader{early, std::thread([&] {
        for (;;) {
            typename Early::Job j;
            {
                std::unique_lock<std::mutex> g(early.m);
                early.cv.wait(g, [&] { return !early.jobs.empty() || early.closed; });
                if (early.jobs.empty()) return;
                j = early.jobs.front();
                early.jobs.pop_front();
            }
            if (early.rc != MZ_OK) continue;  // (only this thread writes rc)
            const int rc = B::h2d_worker(s->ctx, s->device, const_cast<void *>(j.dst), j.src, j.bytes);
            if (rc != MZ_OK) {
                early.rc = rc;
                early.msg = mz::last_error();
            }
        }
    })};
    int rc;
#define RESERVE(field, cnt) if ((rc = dev_reserve<B>(s, d.field, (size_t)(cnt))) != MZ_OK) return rc
    RESERVE(trip_time, n->n_trips);
    RESERVE(trip_length, n->n_trips);
    early_up(d.trip_time, n->trip_time, sizeof(*n->trip_time) * (size_t)n->n_trips);
    early_up(d.trip_length, n->trip_length, sizeof(*n->trip_length) * (size_t)n->n_trips);
    // ---- trips per origin (ODaction execution order) and per OD pair
    std::vector<int> otrip_off(n->n_origins + 1, 0);
    RawVector<int> otrips, trip_slot;  // (the per-trip vectors are sized by their threads)
    RESERVE(origin_trip_off, otrip_off.size());
    RESERVE(origin_trips, n->n_trips);
    RESERVE(trip_slot, n->n_trips);
    Joiner th_origin_trips{std::thread([&] {
        group_trips(n->n_trips, n->trip_origin, n->n_origins, otrip_off, otrips, &trip_slot);
        s->h_otrip_off = otrip_off;
        early_up(d.origin_trip_off, otrip_off.data(), sizeof(int) * otrip_off.size());
        early_up(d.origin_trips, otrips.data(), sizeof(int) * otrips.size());
        early_up(d.trip_slot, trip_slot.data(), sizeof(int) * trip_slot.size());
    })};
    // ---- exit-log offsets, routes with a terminator after each route
    std::vector<int> route_term;
    RawVector<int> trip_rpos0;
    const size_t n_route_term = (size_t)n->route_off[n->n_routes] + (size_t)n->n_routes;
    RESERVE(route_term, n_route_term);
    RESERVE(trip_rpos0, n->n_trips);
    RESERVE(trip_log_off, (size_t)n->n_trips + 1);
    Joiner th_log_off{std::thread([&] {
        s->h_trip_route = vec(n->trip_route, (size_t)n->n_trips);
        s->h_trip_log_off.assign(n->n_trips + 1, 0);
        for (int v = 0; v < n->n_trips; ++v)
            s->h_trip_log_off[v + 1] = s->h_trip_log_off[v] + (n->route_off[n->trip_route[v] + 1] - n->route_off[n->trip_route[v]]);
        s->total_log = s->h_trip_log_off[n->n_trips];
        early_up(d.trip_log_off, s->h_trip_log_off.data(), sizeof(s->h_trip_log_off[0]) * s->h_trip_log_off.size());
    })};
    Joiner th_routes{std::thread([&] {
        s->h_route_off = vec(n->route_off, (size_t)n->n_routes + 1);
        s->h_route_links = vec(n->route_links, (size_t)n->route_off[n->n_routes]);
        route_term.reserve(n_route_term);
        std::vector<int> route_term_off(n->n_routes);
        for (int r = 0; r < n->n_routes; ++r) {
            route_term_off[r] = (int)route_term.size();
            for (int i = n->route_off[r]; i < n->route_off[r + 1]; ++i) route_term.push_back(n->route_links[i]);
            route_term.push_back(-1);
        }
        early_up(d.route_term, route_term.data(), sizeof(int) * std::min(n_route_term, route_term.size()));
        trip_rpos0.resize(n->n_trips);
        for (int v = 0; v < n->n_trips; ++v) trip_rpos0[v] = route_term_off[n->trip_route[v]];
        early_up(d.trip_rpos0, trip_rpos0.data(), sizeof(int) * trip_rpos0.size());
    })};
    std::vector<int> od_off(n->n_odpairs + 1, 0);
    RawVector<int> od_trips;
    RESERVE(od_off, od_off.size());
    RESERVE(od_trips, n->n_trips);
    Joiner th_od_trips{std::thread([&] {  // (joined where the ODactions are built)
        group_trips(n->n_trips, n->trip_od, n->n_odpairs, od_off, od_trips, (RawVector<int> *)nullptr);
        early_up(d.od_off, od_off.data(), sizeof(int) * od_off.size());
        early_up(d.od_trips, od_trips.data(), sizeof(int) * od_trips.size());
    })};
    // ---- node neighbours: nodes sharing a link
    std::vector<int> nb_off, nb;
    Joiner th_neighbours{std::thread([&] { node_neighbours(n, nb_off, nb); })};
    // ---- expected load of a node from the demand (balances the regions): every trip crosses the down node of each link of
    // its route once
    std::vector<double> route_w(N, 0.0);
    Joiner th_weights{std::thread([&] {
        std::vector<int> route_trips(std::max(1, n->n_routes), 0);
        for (int v = 0; v < n->n_trips; ++v) route_trips[n->trip_route[v]]++;
        for (int r = 0; r < n->n_routes; ++r)
            if (route_trips[r])
                for (int i = n->route_off[r]; i < n->route_off[r + 1]; ++i)
                    route_w[n->link_out_node[n->route_links[i]]] += 0.02 * route_trips[r];
    })};
    // ---- packed static records of links and turnings
    std::vector<LinkStat> lstat;
    std::vector<TurnStat> turn;
    RESERVE(lstat, L);
    RESERVE(turn, std::max(1, n->n_turnings));
#undef RESERVE
    std::vector<SdRec> sd(n->n_sdfuncs);
    RawVector<SrvStat> srv;
    Joiner th_records{std::thread([&] {
        for (int i = 0; i < n->n_sdfuncs; ++i)
            sd[i] = SdRec{n->sd_vfree[i], n->sd_vmin[i], n->sd_romax[i], n->sd_romin[i], n->sd_alpha[i], n->sd_beta[i],
                          n->sd_type[i], 0, 0.0};
        srv.resize(std::max(1, n->n_servers));
        srv[0] = SrvStat{0.0, 0.0, 0.0, 0, -1};
        for (int i = 0; i < n->n_servers; ++i) srv[i] = SrvStat{n->srv_mu[i], n->srv_sd[i], n->srv_delay[i], n->srv_type[i], -1};
        lstat.resize(L);
        turn.resize(std::max(1, n->n_turnings));
        for (int i = 0; i < L; ++i)
            lstat[i] = LinkStat{qoff[i], qcap[i], maxcap[i], n->link_lanes[i], n->link_length[i], n->link_sdfunc[i], freeflow[i]};
        for (int k = 0; k < n->n_turnings; ++k)
            turn[k] = TurnStat{n->turn_in[k], n->turn_out[k], n->turn_server[k], n->turn_lookback[k]};
        early_up(d.lstat, lstat.data(), sizeof(LinkStat) * lstat.size());
        early_up(d.turn, turn.data(), sizeof(TurnStat) * turn.size());
    })};
    // on an early return the uploader must stop before the vectors above go away (declared last = destroyed first)
    struct UpGuard {
        Uploader &u;
        ~UpGuard() { u.finish(); }
    } up_guard{th_uploader};

    tick("trips");
    // ---- actions in Network::init order with their first booking (B/network.cpp:7657-7731, B/turning.cpp:207-216)
    struct HAct { int node, kind, idx, aux, sv; double t, tp; int sub; };
    RawVector<HAct> acts;
    double initvalue = 0.1;
    // Traffic signals (B/trafficsignal.cpp): a turning that some Stage regulates starts red (Stage::add_turning) — its
    // TurnActions do nothing at init and are not booked. Each of them gets MZ_SIG_CHAINS event-chain slots instead of one
    // pending event: Turning::activate books the action again while an older booking may still be pending.
    std::vector<int> turn_ctrl(std::max(1, n->n_turnings), -1);  // control that regulates the turning, -1 = unsignalised
    for (int c = 0; c < n->n_sig_controls; ++c)
        for (int p = n->sigc_plan_off[c]; p < n->sigc_plan_off[c + 1]; ++p)
            for (int st = n->sigp_stage_off[p]; st < n->sigp_stage_off[p + 1]; ++st)
                for (int i = n->sigs_turn_off[st]; i < n->sigs_turn_off[st + 1]; ++i) {
                    const int k = n->sigs_turns[i];
                    MZ_REQUIRE(k >= 0 && k < n->n_turnings, MZ_EINVAL, "signal stage %d: turning out of range", st);
                    MZ_REQUIRE(turn_ctrl[k] < 0 || turn_ctrl[k] == c, MZ_ELIMIT, "turning %d is regulated by two signal controls", k);
                    MZ_REQUIRE(n->turn_node[k] == n->turn_node[n->sigs_turns[n->sigs_turn_off[n->sigp_stage_off[n->sigc_plan_off[c]]]]],
                               MZ_ELIMIT, "signal control %d regulates turnings of several nodes", c);
                    turn_ctrl[k] = c;
                }
    std::vector<int> turn_act0(std::max(1, n->n_turnings), 0), turn_nact(std::max(1, n->n_turnings), 0);
    {
        // Network::init's clock advances by repeated addition over the turnings (the order matters); the records
        // themselves are independent of each other and are written by a few threads
        std::vector<double> turn_t0(std::max(1, n->n_turnings));
        size_t na = 0;
        for (int k = 0; k < n->n_turnings; ++k) {
            const int cnt = std::min(n->link_lanes[n->turn_in[k]], n->link_lanes[n->turn_out[k]]);
            turn_act0[k] = (int)na;
            turn_nact[k] = cnt;
            turn_t0[k] = initvalue;
            na += (size_t)std::max(0, cnt) * (turn_ctrl[k] < 0 ? 1 : MZ_SIG_CHAINS);
            initvalue += 0.00001;
        }
        MZ_REQUIRE(na < (size_t)INT32_MAX / 2, MZ_ELIMIT, "more than 2^30 turn actions");
        acts.reserve(na + (size_t)n->n_odpairs + (size_t)n->n_dests * 8 + (size_t)n->n_sig_controls + (size_t)n->n_incidents + 4096);
        acts.resize(na);
        auto fill = [&](int k0, int k1) {
            for (int k = k0; k < k1; ++k) {
                double t = turn_t0[k];
                size_t at = (size_t)turn_act0[k];
                for (int j = 0; j < turn_nact[k]; ++j) {
                    if (turn_ctrl[k] < 0) {
                        // executed at init on an empty network: out-link not full, in-link empty ⇒ nexttime = t + freeflowtime
                        acts[at++] = HAct{n->turn_node[k], ACT_TURN, k, 0, -1, t + freeflow[n->turn_in[k]], t, 0};
                    } else {
                        for (int q = 0; q < MZ_SIG_CHAINS; ++q)
                            acts[at++] = HAct{n->turn_node[k], ACT_TURN, k, 0, -1, INFINITY, INFINITY, 0};
                    }
                    t += 0.00000003;
                }
            }
        };
        const int n_thr = n->n_turnings > (1 << 18) ? 4 : 1;
        std::vector<std::thread> th;
        for (int k = 1; k < n_thr; ++k)
            th.emplace_back(fill, (int)((long long)n->n_turnings * k / n_thr), (int)((long long)n->n_turnings * (k + 1) / n_thr));
        fill(0, (int)((long long)n->n_turnings / n_thr));
        for (std::thread &t : th) t.join();
    }
    tick("actions: turnings");
    // Signal controls: SignalControl::execute at init, 0.000001 apart (B/network.cpp:7666-7670). All signal events are
    // independent of the traffic state: the control's event sequence — times, FIFO order among themselves, the turnings each
    // event stops and activates — is computed here by running the Control / Plan / Stage automaton on its own, and becomes
    // ONE action of the junction's node that walks the list (sim_core.h: ACT_SIGNAL).
    std::vector<SigEv> sig_ev;
    std::vector<int> sig_ops, sigc_ev0(1, 0);
    std::vector<unsigned char> turn_active0((size_t)std::max(1, n->n_turnings), 1);
    for (int k = 0; k < n->n_turnings; ++k)
        if (turn_ctrl[k] >= 0) turn_active0[k] = 0;
    long long n_sig_init_events = 0;
    for (int c = 0; c < n->n_sig_controls; ++c) {
        struct Book { double t; long long seq; int kind, idx, parent; };  // kind 0 control, 1 plan, 2 stage
        std::vector<Book> heap;
        long long seq = 0;
        const int p0 = n->sigc_plan_off[c], p1 = n->sigc_plan_off[c + 1];
        MZ_REQUIRE(p1 > p0, MZ_EINVAL, "signal control %d has no plan", c);
        std::vector<char> plan_active(p1 - p0, 0), stage_active(std::max(1, n->n_sig_stages), 0);
        std::vector<int> plan_cur(p1 - p0, 0);
        int ctrl_cur = p0;
        bool ctrl_active = false;
        std::vector<int> ops;
        int cur_event = -1;
        auto book = [&](double t, int kind, int idx) { heap.push_back(Book{t, seq++, kind, idx, cur_event}); };
        auto stage_stop = [&](int st) {
            stage_active[st] = 0;
            for (int i = n->sigs_turn_off[st]; i < n->sigs_turn_off[st + 1]; ++i) ops.push_back(n->sigs_turns[i] << 1);
        };
        auto stage_execute = [&](int st, double time) {
            if (!stage_active[st]) {
                for (int i = n->sigs_turn_off[st]; i < n->sigs_turn_off[st + 1]; ++i) ops.push_back((n->sigs_turns[i] << 1) | 1);
                stage_active[st] = 1;
                book(time + n->sigs_duration[st], 2, st);
            } else {
                stage_stop(st);
            }
        };
        auto plan_execute = [&](int p, double time) {
            const int s0 = n->sigp_stage_off[p], s1 = n->sigp_stage_off[p + 1];
            if (!plan_active[p - p0]) {
                plan_cur[p - p0] = s0;
                plan_active[p - p0] = 1;
            }
            if (time < n->sigp_stop[p]) {
                const int cur = plan_cur[p - p0];
                stage_execute(cur, time);
                const double next_stage_time = time + n->sigs_duration[cur];
                plan_cur[p - p0] = cur + 1 == s1 ? s0 : cur + 1;
                book(next_stage_time, 1, p);
            } else {
                plan_active[p - p0] = 0;
                for (int st = s0; st < s1; ++st) stage_stop(st);
            }
        };
        auto control_execute = [&](double time) {
            if (!ctrl_active) {
                ctrl_cur = p0;
                ctrl_active = true;
            }
            plan_execute(ctrl_cur, time);
            ++ctrl_cur;
            if (ctrl_cur < p1) book(n->sigp_start[ctrl_cur], 0, c);
        };
        for (int p = p0; p < p1; ++p) MZ_REQUIRE(n->sigp_stage_off[p + 1] > n->sigp_stage_off[p], MZ_EINVAL, "signal plan %d has no stage", p);
        const double t_init = initvalue;
        control_execute(t_init);
        // the init execution activates the first stage on the empty network: every TurnAction of its turnings executes at
        // t_init (+3e-8 per action) and books its first chain at time + freeflowtime; sub = order of the bookings
        int node_c = -1, sub = 0;
        for (int op : ops) {
            const int k = op >> 1;
            node_c = n->turn_node[k];
            if (!(op & 1)) {
                turn_active0[k] = 0;
                continue;
            }
            turn_active0[k] = 1;
            double t = t_init;
            for (int j = 0; j < turn_nact[k]; ++j) {
                HAct &slot = acts[(size_t)turn_act0[k] + (size_t)j * MZ_SIG_CHAINS];
                slot.t = t + freeflow[n->turn_in[k]];
                slot.tp = t_init;
                slot.sub = ++sub;
                t += 0.00000003;
                ++n_sig_init_events;
            }
        }
        const int init_end_sub = ++sub;
        ops.clear();
        if (node_c < 0) node_c = n->turn_node[n->sigs_turns[n->sigs_turn_off[n->sigp_stage_off[p0]]]];
        // the control's events in execution order, up to and including the first one at or beyond the horizon
        const size_t ev0 = sig_ev.size();
        std::vector<double> ev_time;  // time of each event of this control (for the children's tp)
        while (!heap.empty()) {
            size_t best = 0;
            for (size_t i = 1; i < heap.size(); ++i)
                if (heap[i].t < heap[best].t || (heap[i].t == heap[best].t && heap[i].seq < heap[best].seq)) best = i;
            const Book b = heap[best];
            heap.erase(heap.begin() + (long)best);
            cur_event = (int)(sig_ev.size() - ev0);
            if (b.kind == 0) control_execute(b.t);
            else if (b.kind == 1) plan_execute(b.idx, b.t);
            else stage_execute(b.idx, b.t);
            SigEv e;
            e.t = b.t;
            e.tp = b.parent < 0 ? t_init : ev_time[(size_t)b.parent];
            e.op_off = (int)sig_ops.size();
            e.n_ops = (int)ops.size();
            e.parent = b.parent < 0 ? -1 : (int)ev0 + b.parent;
            e.sub0 = init_end_sub;
            sig_ops.insert(sig_ops.end(), ops.begin(), ops.end());
            ops.clear();
            sig_ev.push_back(e);
            ev_time.push_back(b.t);
            if (!(b.t < n->runtime)) break;
        }
        sigc_ev0.push_back((int)sig_ev.size());
        const bool any = sig_ev.size() > ev0;
        acts.push_back(HAct{node_c, ACT_SIGNAL, c, 0, -1, any ? sig_ev[ev0].t : INFINITY, any ? sig_ev[ev0].tp : INFINITY,
                            init_end_sub});
        ++n_sig_init_events;
        initvalue += 0.000001;
    }
    // the transit layer's share of Network::init: one 0.00001 step per Busline and per active ODstops pair (B/network.cpp:7672-7694)
    for (int i = 0; i < n->n_transit_init_steps; ++i) initvalue += 0.00001;
    const size_t n_turn_acts = acts.size();
    th_od_trips.t.join();
    // The trips of one trip_od value are ONE event chain: a trip is booked when the chain's previous trip executes, never
    // before. A table that files somebody else's vehicles into a chain (a bus line's dispatches into the OD pair that shares
    // its origin) breaks that, and trips of equal time would then overtake each other in the origin's virtual queue — so it
    // is refused here, by a pass over the chains that runs beside the rest of the build.
    int chain_bad_trip = -1;
    Joiner th_chain_check{std::thread([&] {
        for (int k = 0; k < n->n_odpairs && chain_bad_trip < 0; ++k)
            for (int i = od_off[k] + 1; i < od_off[k + 1]; ++i)
                if (n->trip_prev_time[od_trips[i]] < n->trip_time[od_trips[i - 1]]) {
                    chain_bad_trip = od_trips[i];
                    break;
                }
    })};
    for (int k = 0; k < n->n_odpairs; ++k) {
        // ODpair::execute booked the pair's first ODaction at init (B/od.cpp:355-368): an action like any other
        if (od_off[k] < od_off[k + 1]) {
            const int v0 = od_trips[od_off[k]];
            acts.push_back(HAct{n->origin_node[n->trip_origin[v0]], ACT_OD, k, n->trip_origin[v0], -1, n->trip_time[v0],
                                n->trip_prev_time[v0], 0});
        }
        initvalue += 0.00001;
    }
    const size_t n_od_acts = acts.size() - n_turn_acts;
    // in-links per node, ascending link id
    std::vector<int> in_off(N + 1, 0), in_links(L);
    for (int l = 0; l < L; ++l) in_off[n->link_out_node[l] + 1]++;
    for (int i = 0; i < N; ++i) in_off[i + 1] += in_off[i];
    {
        std::vector<int> c(N, 0);
        for (int l = 0; l < L; ++l) in_links[in_off[n->link_out_node[l]] + c[n->link_out_node[l]]++] = l;
    }
    for (int k = 0; k < n->n_dests; ++k) {
        const int dn = n->dest_node[k];
        for (int li = in_off[dn + 1] - 1; li >= in_off[dn]; --li) {  // dactions.insert(begin()) over ascending link ids ⇒ descending order
            const int l = in_links[li];
            for (int j = 0; j < n->link_lanes[l]; ++j)
                acts.push_back(HAct{n->dest_node[k], ACT_DEST, l, n->n_servers + k, n->dest_server[k],
                                    initvalue + freeflow[l], initvalue, 0});  // empty link ⇒ next_action = t + freeflowtime
        }
        initvalue += 0.00001;
    }
    // incidents without information broadcast: one action per incident at the link's downstream node. Incident::Incident
    // books the start while the incident file is read, which executemaster does right AFTER Network::init
    // (B/network.cpp:7380-7383): the booking follows all of init's (parent time = the init clock where init left it) and
    // precedes everything booked during the run.
    const size_t n_inc_acts = (size_t)n->n_incidents;
    for (int i = 0; i < n->n_incidents; ++i)
        acts.push_back(HAct{n->link_out_node[n->inc_link[i]], ACT_INCIDENT, n->inc_link[i], i, -1, n->inc_start[i], initvalue, 0});
    // transit overlay hooks: one ACT_STOP action per link that carries a stop, at the link's downstream node; it has no
    // event until the host books a departure (mz_sim_stop_departures)
    size_t n_stop_acts = 0;
    std::vector<size_t> slink_act_raw;
    s->h_slink_link.clear();
    if (n->trip_stop_off) {
        const int n_stops = n->trip_stop_off[n->n_trips];
        std::vector<int> sl(n->stop_link, n->stop_link + n_stops);
        std::sort(sl.begin(), sl.end());
        sl.erase(std::unique(sl.begin(), sl.end()), sl.end());
        s->h_slink_link = sl;
        for (size_t i = 0; i < sl.size(); ++i) {
            slink_act_raw.push_back(acts.size());
            acts.push_back(HAct{n->link_out_node[sl[i]], ACT_STOP, sl[i], (int)i, -1, INFINITY, INFINITY, 0});
        }
        n_stop_acts = sl.size();
    }
    // turnactions and dactions execute once inside init(); red turnings' actions do not (their chain slots are not events)
    {
        long long red_slots = 0;
        for (int k = 0; k < n->n_turnings; ++k)
            if (turn_ctrl[k] >= 0) red_slots += (long long)turn_nact[k] * MZ_SIG_CHAINS;
        s->n_init_events = (long long)(acts.size() - n_od_acts - n_inc_acts - n_stop_acts) - red_slots - n->n_sig_controls + n_sig_init_events;
    }
    tick("actions: signals, ODs, sinks");
    // cross-node sharing of a stochastic server would serialise the whole network through one RNG stream;
    // checked beside the grouping below
    int shared_sv = -1;
    Joiner th_server_check{std::thread([&] {
        std::vector<int> owner((size_t)std::max(1, n->n_servers), -1);  // dense server index -> first node using it
        for (const HAct &a : acts) {
            const int sv = a.kind == ACT_TURN ? n->turn_server[a.idx] : a.sv;
            if (sv < 0 || (n->srv_type[sv] != 1 && n->srv_type[sv] != 4)) continue;
            if (owner[sv] < 0) owner[sv] = a.node;
            else if (owner[sv] != a.node) {
                shared_sv = sv;
                return;
            }
        }
    })};
    // group by node (counting sort: stable, keeps init order inside a node)
    RawVector<int> order;
    order.resize(acts.size());
    std::vector<int> act_off;  // first action of every node
    {
        std::vector<int> pos(N + 1, 0);
        for (const HAct &a : acts) pos[a.node + 1]++;
        for (int i = 0; i < N; ++i) pos[i + 1] += pos[i];
        act_off = pos;
        for (size_t i = 0; i < acts.size(); ++i) order[pos[acts[i].node]++] = (int)i;
    }
    th_server_check.t.join();
    MZ_REQUIRE(shared_sv < 0, MZ_ELIMIT,
               "stochastic server %d is shared by several nodes: its draws would have to follow the global "
               "event order (SURVEY.md H5); give each node its own server or use deterministic servers", shared_sv);
    tick("actions: order");
    const int A = (int)acts.size();
    s->n_acts = A;
    d.n_acts = A;
    RawVector<ActStat> astat;
    astat.resize(std::max(1, A));
    astat[0] = ActStat{0, 0, 0, 0, 0, 0, 0, 0};
    s->h_adyn.resize(A);
    s->h_act_node.resize(A);
    {
        // the records in node order: a gather over the 3 M+ actions of a large network, split over a few threads
        auto fill = [&](int i0, int i1) {
            for (int i = i0; i < i1; ++i) {
                const HAct &a = acts[order[i]];
                // sink: aux = slot of its server seed, aux2 = server index (or -1 for the node's own default server);
                // ODaction: aux = origin index
                astat[i] = ActStat{a.kind, a.idx, a.aux, a.sv, 0, 0, 0, 0};
                if (a.kind == ACT_TURN) {
                    astat[i].in = n->turn_in[a.idx];
                    astat[i].out = n->turn_out[a.idx];
                    astat[i].server = n->turn_server[a.idx];
                    astat[i].lookback = n->turn_lookback[a.idx];
                }
                s->h_adyn[i] = ActDyn{a.t, a.tp, a.sub, 0, 0.0};
                s->h_act_node[i] = a.node;
            }
        };
        const int n_thr = A > (1 << 18) ? 6 : 1;
        std::vector<std::thread> th;
        for (int k = 1; k < n_thr; ++k) th.emplace_back(fill, (int)((long long)A * k / n_thr), (int)((long long)A * (k + 1) / n_thr));
        fill(0, (int)((long long)A / n_thr));
        for (std::thread &t : th) t.join();
    }
    // where a signalised turning's chain slots ended up (contiguous: same node, stable grouping) and which nodes own signals
    std::vector<int> turn_slot0(std::max(1, n->n_turnings), -1);
    std::vector<unsigned char> node_sig((size_t)N, 0);
    s->h_slink_act.resize(slink_act_raw.size());
    if (n->n_sig_controls > 0 || !slink_act_raw.empty()) {  // (the inverse permutation is only needed for these two features)
        std::vector<int> final_of(acts.size());
        for (int i = 0; i < A; ++i) final_of[order[i]] = i;
        for (int k = 0; k < n->n_turnings; ++k)
            if (turn_ctrl[k] >= 0) {
                turn_slot0[k] = final_of[turn_act0[k]];
                node_sig[n->turn_node[k]] = 1;
            }
        for (size_t i = 0; i < slink_act_raw.size(); ++i) s->h_slink_act[i] = final_of[slink_act_raw[i]];
    }
    s->h_turn_active0 = turn_active0;

    tick("actions");
    th_neighbours.t.join();
    s->h_nb_off = nb_off;
    s->h_nb = nb;
    s->h_link_in = vec(n->link_in_node, (size_t)L);
    s->h_link_out = vec(n->link_out_node, (size_t)L);
    tick("neighbours");
    th_origin_trips.t.join();
    th_routes.t.join();
    th_log_off.t.join();
    MZ_REQUIRE(s->total_log < (long long)INT32_MAX, MZ_ELIMIT, "exit log exceeds 2^31 records");

    tick("routes/log");
    // ---- ownership: a node belongs to its rank, a link to the rank of its downstream node
    s->h_node_home.assign(N, 0);
    if (s->n_ranks > 1)
        for (int i = 0; i < N; ++i) s->h_node_home[i] = (unsigned char)s->h_node_rank[i];
    std::vector<unsigned char> link_home(L);
    for (int l = 0; l < L; ++l) link_home[l] = s->h_node_home[n->link_out_node[l]];
    std::vector<int> my_nodes;
    for (int i = 0; i < N; ++i)
        if (s->h_node_home[i] == s->rank) my_nodes.push_back(i);
    // ---- regions (one per thread block of the node-visit kernel): the bisection of this rank's nodes runs beside the rest
    // of the build; sim_set_regions picks the result up. Node weight = its actions plus the traversals it will perform.
    th_weights.t.join();
    s->h_my_nodes = my_nodes;
    s->h_node_weight.assign(N, 1.0);
    for (int i = 0; i < N; ++i) s->h_node_weight[i] += 0.25 * (act_off[i + 1] - act_off[i]) + route_w[i];
    d.n_my = (int)my_nodes.size();
    s->region_cap = std::max(1, std::min(d.n_my, 1 << 16));
    s->pre_regions = std::max(1, std::min(B::default_regions(s), std::min(std::max(1, d.n_my), s->region_cap)));
    Joiner th_regions{std::thread([&] {
        s->h_region_pre.assign(N, -1);
        std::vector<int> ids = s->h_my_nodes, member(N, 0), seen(N, 0);
        int stamp = 0;
        region_bisect(s->h_nb_off, s->h_nb, s->h_node_weight, ids, 0, s->pre_regions, s->h_region_pre, member, seen, stamp);
    })};
    // replication across the cut: a cut link is held by the ranks of its two end nodes, a boundary node's published
    // state by the ranks of its neighbours; the owner writes every copy, everybody reads its own
    std::vector<unsigned char> link_peer(L, 255), node_peers(N, 0);
    for (int l = 0; l < L; ++l) {
        const int a = n->link_in_node[l], b = n->link_out_node[l];
        const int ra = s->h_node_home[a], rb = s->h_node_home[b];
        if (ra == rb) continue;
        if (ra == s->rank) link_peer[l] = (unsigned char)rb;
        else if (rb == s->rank) link_peer[l] = (unsigned char)ra;
        node_peers[a] |= (unsigned char)(1u << rb);
        node_peers[b] |= (unsigned char)(1u << ra);
    }

    // ---- initial node keys
    s->h_key_t.assign(N, INFINITY);
    s->h_key_tp.assign(N, INFINITY);
    {
        auto keys = [&](int i0, int i1) {
            for (int i = i0; i < i1; ++i) {
                Key best{INFINITY, INFINITY, 0x7fffffff, 0x7fffffff};
                for (int a = act_off[i]; a < act_off[i + 1]; ++a) {
                    Key k{s->h_adyn[a].t, s->h_adyn[a].tp, s->h_adyn[a].sub, a};
                    if (key_less(k, best)) best = k;
                }
                s->h_key_t[i] = best.t;
                s->h_key_tp[i] = best.tp;
            }
        };
        const int n_thr = N > (1 << 16) ? 4 : 1;
        std::vector<std::thread> th;
        for (int k = 1; k < n_thr; ++k) th.emplace_back(keys, (int)((long long)N * k / n_thr), (int)((long long)N * (k + 1) / n_thr));
        keys(0, (int)((long long)N / n_thr));
        for (std::thread &t : th) t.join();
    }

    tick("keys");
    // ---- packed static records (links and turnings: th_records)
    th_records.t.join();
    // server-rate changes: per server in time order (stable: equal times in file order), folded into the rates in force
    // from each change on ("mu > 0 ? set_mu", "sd > 0 ? set_sd", B/server.cpp:161-168), closed by a +inf entry
    std::vector<SrvChg> srv_chg(1, SrvChg{INFINITY, 0.0, 0.0, 0.0});
    if (n->n_rate_changes > 0) {
        MZ_REQUIRE(n->chg_server && n->chg_time && n->chg_mu && n->chg_sd, MZ_EINVAL, "rate-change arrays missing");
        std::vector<int> order(n->n_rate_changes);
        for (int i = 0; i < n->n_rate_changes; ++i) {
            MZ_REQUIRE(n->chg_server[i] >= 0 && n->chg_server[i] < n->n_servers, MZ_EINVAL, "rate change %d: server out of range", i);
            order[i] = i;
        }
        std::stable_sort(order.begin(), order.end(), [&](int a, int b) {
            if (n->chg_server[a] != n->chg_server[b]) return n->chg_server[a] < n->chg_server[b];
            return n->chg_time[a] < n->chg_time[b];
        });
        srv_chg.clear();
        for (size_t k = 0; k < order.size();) {
            const int sv = n->chg_server[order[k]];
            srv[sv].chg = (int)srv_chg.size();
            double mu = n->srv_mu[sv], sd = n->srv_sd[sv];
            for (; k < order.size() && n->chg_server[order[k]] == sv; ++k) {
                const int c = order[k];
                if (n->chg_mu[c] > 0.0) mu = n->chg_mu[c];
                if (n->chg_sd[c] > 0.0) sd = n->chg_sd[c];
                srv_chg.push_back(SrvChg{n->chg_time[c], mu, sd, 0.0});
            }
            srv_chg.push_back(SrvChg{INFINITY, mu, sd, 0.0});
        }
    }

    tick("records");
    // ---- upload
    d.n_ranks = s->n_ranks;
    d.rank = s->rank;
    // (lstat, turn, route_term and the per-trip tables are already on their way: early_up above)
    th_chain_check.t.join();
    MZ_REQUIRE(chain_bad_trip < 0, MZ_EINVAL,
               "trip %d (trip_od %d) was booked before the previous trip of its chain ran (trip_prev_time %.17g): the trips of one "
               "trip_od value must form one event chain — give a vehicle source of its own (a bus line) its own trip_od",
               chain_bad_trip, n->trip_od[chain_bad_trip], n->trip_prev_time[chain_bad_trip]);
    th_uploader.finish();
    if (early.rc != MZ_OK) {
        mz::set_error("%s", early.msg.c_str());
        return early.rc;
    }
#define UP(field, v) if ((rc = dev_upload<B>(s, d.field, v)) != MZ_OK) return rc
    UP(link_hist, s->h_link_hist);
    UP(sd, sd);
    UP(srv, srv);
    UP(srv_chg, srv_chg);
    if (n->n_incidents > 0) {
        UP(inc_stop, vec(n->inc_stop, (size_t)n->n_incidents));
        UP(inc_sdfunc, vec(n->inc_sdfunc, (size_t)n->n_incidents));
    }
    if (n->n_sig_controls > 0) {
        UP(sig_ev, sig_ev);
        UP(sig_ops, sig_ops);
        UP(sigc_ev0, sigc_ev0);
        UP(turn_slot0, turn_slot0);
        UP(turn_nact, turn_nact);
        UP(node_sig, node_sig);
    }
    if (n->trip_stop_off) {
        const size_t ns = (size_t)n->trip_stop_off[n->n_trips];
        s->n_stops = (int)ns;
        s->n_slinks = (int)s->h_slink_link.size();
        s->h_trip_stop_off = vec(n->trip_stop_off, (size_t)n->n_trips + 1);
        s->h_stop_link = vec(n->stop_link, ns);
        UP(trip_stop_off, s->h_trip_stop_off);
        UP(stop_link, s->h_stop_link);
        UP(stop_pos, vec(n->stop_pos, ns));
    }
    UP(astat, astat);
    UP(node_act_off, act_off);
    UP(node_nb_off, nb_off);
    UP(link_home, link_home);
    UP(node_home, s->h_node_home);
    UP(link_peer, link_peer);
    UP(node_peers, node_peers);
#undef UP
    tick("upload (rest)");
#define AL(field, cnt) if ((rc = dev_alloc<B>(s, d.field, (size_t)(cnt))) != MZ_OK) return rc
    AL(lacc, L); AL(avgtimes, (size_t)L * P);
    AL(adyn, A); AL(turn_blocked, n->n_turnings);
    if (n->n_sig_controls > 0) {
        AL(turn_active, n->n_turnings);
        AL(sig_next, n->n_sig_controls);
        AL(sig_endsub, sig_ev.size() + 1);
        s->n_sig_events = (long long)sig_ev.size();
    }
    s->n_servers_total = n->n_servers + n->n_dests;
    AL(srv_seed, s->n_servers_total);
    AL(od_next, n->n_odpairs); AL(origin_vq_head, n->n_origins); AL(trip_parked, n->n_trips);
    AL(arrival, n->n_trips); AL(exit_log, s->total_log);
    AL(counters, 64);
    if (n->trip_stop_off) {
        AL(bus_next, s->n_stops); AL(bus_at, s->n_stops); AL(stop_out, s->n_stops); AL(slink_cur, s->n_slinks);
        if ((rc = dev_alloc<B>(s, s->d_dep_off, (size_t)s->n_slinks + 1)) != MZ_OK) return rc;
        if ((rc = dev_alloc<B>(s, s->d_dep_veh, (size_t)s->n_stops)) != MZ_OK) return rc;
        if ((rc = dev_alloc<B>(s, s->d_dep_t, (size_t)s->n_stops)) != MZ_OK) return rc;
        d.dep_off = s->d_dep_off;
        d.dep_veh = s->d_dep_veh;
        d.dep_t = s->d_dep_t;
    }
    s->horizon = n->runtime;
#undef AL
    if ((rc = dev_alloc<B>(s, s->d_my_nodes, (size_t)d.n_my)) != MZ_OK) return rc;
    if ((rc = dev_alloc<B>(s, s->d_region_off, (size_t)s->region_cap + 1)) != MZ_OK) return rc;
    if ((rc = dev_alloc<B>(s, s->d_node_nb, nb.size())) != MZ_OK) return rc;
    if ((rc = dev_alloc<B>(s, s->d_nb_slot, nb.size())) != MZ_OK) return rc;
    if ((rc = dev_alloc<B>(s, s->d_mb_off, (size_t)s->region_cap + 1)) != MZ_OK) return rc;
    if ((rc = dev_alloc<B>(s, s->d_poll_off, (size_t)s->region_cap + 1)) != MZ_OK) return rc;
    if ((rc = dev_alloc<B>(s, s->d_poll_pos, (size_t)d.n_my)) != MZ_OK) return rc;
    s->mailbox_words = (size_t)d.n_my / 32 + (size_t)s->region_cap + 2;
    if ((rc = dev_alloc<B>(s, d.mailbox, s->mailbox_words)) != MZ_OK) return rc;
    d.nb_slot = s->d_nb_slot;
    d.mb_off = s->d_mb_off;
    d.poll_off = s->d_poll_off;
    d.poll_pos = s->d_poll_pos;
    if ((rc = dev_alloc<B>(s, s->d_node_cls, (size_t)N)) != MZ_OK) return rc;
    if ((rc = dev_alloc<B>(s, s->d_link_cls, (size_t)L)) != MZ_OK) return rc;
    d.my_nodes = s->d_my_nodes;
    d.region_off = s->d_region_off;
    d.node_nb = s->d_node_nb;
    d.node_cls = s->d_node_cls;
    d.link_cls = s->d_link_cls;
    tick("alloc");
    // ---- the arena: identical layout on every rank (peers address it by the same offsets)
    {
        size_t off = 0;
        auto take = [&](size_t bytes) {
            const size_t o = off;
            off += (bytes + 255) / 256 * 256;
            return o;
        };
        d.off_hot = take(sizeof(LinkHot) * (size_t)L);
        d.off_slab = take(sizeof(QEnt) * (size_t)slab);
        d.off_kt = take(sizeof(double) * (size_t)N);
        d.off_kver = take(sizeof(unsigned) * (size_t)N);
        d.off_kslot[0] = take(sizeof(KeySlot) * (size_t)N);
        d.off_kslot[1] = take(sizeof(KeySlot) * (size_t)N);
        d.off_last = take(sizeof(NodeLast) * (size_t)N);
        s->off_sync = take(sizeof(unsigned long long) * 64);
        s->arena_bytes = off;
        if ((rc = B::arena_alloc(s->ctx, &s->arena, off, s->rank)) != MZ_OK) return rc;
        for (int r = 0; r < MZ_MAXR; ++r) d.arena[r] = nullptr;
        d.arena[s->rank] = (char *)s->arena;
    }
    tick("arena");
    th_regions.t.join();
    if ((rc = sim_set_regions(s, B::default_regions(s))) != MZ_OK) return rc;
    tick("regions");
    const int rc_reset = sim_reset_state(s);
    tick("reset");
    return rc_reset;
}


// Splits this rank's nodes into `n_regions` regions (one per thread block of the node-visit kernel) and derives the sharing
// classes the kernels pick their memory path by (sim_core.h, "memory access"): a node whose neighbours all sit in its own
// region, and a link whose two end nodes do, are touched by ONE SM only. Re-run when the block count changes.
template <class B>
int sim_set_regions(SimHost<B> *s, int n_regions)
{
    SimDev &d = s->dev;
    const int N = s->n_nodes, L = s->n_links, n_my = (int)s->h_my_nodes.size();
    n_regions = std::max(1, std::min(n_regions, std::min(std::max(1, n_my), s->region_cap)));
    if (n_regions == s->n_regions) return MZ_OK;
    if (int rc = B::set_device(s->device)) return rc;
    std::vector<int> region_of;
    if (s->pre_regions == n_regions && (int)s->h_region_pre.size() == N) {
        region_of.swap(s->h_region_pre);  // cut by sim_build's worker thread
    } else {
        region_of.assign(N, -1);
        std::vector<int> ids = s->h_my_nodes, member(N, 0), seen(N, 0);
        int stamp = 0;
        region_bisect(s->h_nb_off, s->h_nb, s->h_node_weight, ids, 0, n_regions, region_of, member, seen, stamp);
    }
    s->pre_regions = 0;
    std::vector<int>().swap(s->h_region_pre);
    // nodes in region order (stable: ascending node id inside a region)
    std::vector<int> region_off(n_regions + 1, 0), my_nodes(n_my);
    for (int v : s->h_my_nodes) region_off[region_of[v] + 1]++;
    int max_region = 0;
    for (int r = 0; r < n_regions; ++r) {
        max_region = std::max(max_region, region_off[r + 1]);
        region_off[r + 1] += region_off[r];
    }
    {
        std::vector<int> pos(region_off.begin(), region_off.end() - 1);
        for (int v : s->h_my_nodes) my_nodes[pos[region_of[v]]++] = v;
    }
    // sharing classes. Nodes of other ranks only ever appear as neighbours across the cut: class 2.
    std::vector<unsigned char> node_cls(N, 2), link_cls(L, 2);
    for (int v : s->h_my_nodes) {
        int c = 0;
        for (int i = s->h_nb_off[v]; i < s->h_nb_off[v + 1]; ++i) {
            const int u = s->h_nb[i];
            c = std::max(c, region_of[u] < 0 ? 2 : (region_of[u] != region_of[v] ? 1 : 0));
        }
        node_cls[v] = (unsigned char)c;
    }
    for (int l = 0; l < L; ++l) {
        const int ra = region_of[s->h_link_in[l]], rb = region_of[s->h_link_out[l]];
        link_cls[l] = (unsigned char)((ra < 0 || rb < 0) ? 2 : (ra != rb ? 1 : 0));
    }
    std::vector<int> nbx(s->h_nb.size());
    for (size_t i = 0; i < nbx.size(); ++i) {
        MZ_REQUIRE(s->h_nb[i] <= MZ_NB_MASK, MZ_ELIMIT, "more than 2^30 nodes");
        nbx[i] = s->h_nb[i] | ((int)node_cls[s->h_nb[i]] << MZ_NB_SHIFT);
    }
    // wake-up tables (sim_core.h: SimDev::nb_slot): position of every node inside its region, mailbox words per region,
    // and per region the nodes that have a neighbour on another rank (nobody can wake those: they are polled)
    std::vector<int> pos_of(N, -1), mb_off(n_regions + 1, 0), poll_off(n_regions + 1, 0), poll_pos;
    for (int r = 0; r < n_regions; ++r) {
        for (int i = region_off[r]; i < region_off[r + 1]; ++i) pos_of[my_nodes[i]] = i - region_off[r];
        mb_off[r + 1] = mb_off[r] + (region_off[r + 1] - region_off[r] + 31) / 32;
    }
    MZ_REQUIRE((size_t)mb_off[n_regions] <= s->mailbox_words, MZ_ELIMIT, "mailbox sizing");
    for (int r = 0; r < n_regions; ++r) {
        for (int i = region_off[r]; i < region_off[r + 1]; ++i)
            if (node_cls[my_nodes[i]] == 2) poll_pos.push_back(i - region_off[r]);
        poll_off[r + 1] = (int)poll_pos.size();
    }
    std::vector<int> nb_slot(s->h_nb.size(), -1);
    for (int v : s->h_my_nodes)
        for (int i = s->h_nb_off[v]; i < s->h_nb_off[v + 1]; ++i) {
            const int u = s->h_nb[i];
            if (region_of[u] < 0) nb_slot[i] = -1;
            else if (region_of[u] == region_of[v]) nb_slot[i] = pos_of[u];
            else nb_slot[i] = -(mb_off[region_of[u]] * 32 + pos_of[u] + 2);
        }
    typename B::Ctx &c = s->ctx;
#define RC(x) if (int rc__ = (x)) return rc__
    if (!nb_slot.empty()) RC(B::h2d(c, s->d_nb_slot, nb_slot.data(), sizeof(int) * nb_slot.size()));
    RC(B::h2d(c, s->d_mb_off, mb_off.data(), sizeof(int) * mb_off.size()));
    RC(B::h2d(c, s->d_poll_off, poll_off.data(), sizeof(int) * poll_off.size()));
    if (!poll_pos.empty()) RC(B::h2d(c, s->d_poll_pos, poll_pos.data(), sizeof(int) * poll_pos.size()));
    if (n_my) RC(B::h2d(c, s->d_my_nodes, my_nodes.data(), sizeof(int) * (size_t)n_my));
    RC(B::h2d(c, s->d_region_off, region_off.data(), sizeof(int) * region_off.size()));
    if (!nbx.empty()) RC(B::h2d(c, s->d_node_nb, nbx.data(), sizeof(int) * nbx.size()));
    RC(B::h2d(c, s->d_node_cls, node_cls.data(), (size_t)N));
    RC(B::h2d(c, s->d_link_cls, link_cls.data(), (size_t)L));
    RC(B::sync(c));
#undef RC
    s->n_regions = n_regions;
    s->max_region = max_region;
    d.n_regions = n_regions;
    return MZ_OK;
}

// (Re)initialises all mutable state: empty queues, Network::init's first bookings, fresh Random seeds.
template <class B>
int sim_reset_state(SimHost<B> *s)
{
    SimDev &d = s->dev;
    if (int rc = B::set_device(s->device)) return rc;
    typename B::Ctx &c = s->ctx;
    const size_t L = (size_t)s->n_links, T = (size_t)std::max(1, s->n_trips), N = (size_t)s->n_nodes;
    char *ar = (char *)s->arena;
#define RC(x) if (int rc__ = (x)) return rc__
    const size_t A = (size_t)s->n_acts;
    if (!s->init_images) {  // first call: build the pristine images on the device
        RC((dev_alloc<B>(s, s->d_adyn0, A)));
        RC((dev_alloc<B>(s, s->d_key_t0, N)));
        RC((dev_alloc<B>(s, s->d_kslot0, N)));
        if (A) RC(B::h2d(c, s->d_adyn0, s->h_adyn.data(), sizeof(ActDyn) * A));
        RC(B::h2d(c, s->d_key_t0, s->h_key_t.data(), sizeof(double) * N));
        std::vector<KeySlot> ks(N);
        for (size_t i = 0; i < N; ++i) ks[i] = KeySlot{s->h_key_t[i], s->h_key_tp[i], 0, 0, 0.0};
        RC(B::h2d(c, s->d_kslot0, ks.data(), sizeof(KeySlot) * N));  // publication 0 lives in slot 0
        if (d.turn_active) {
            RC((dev_alloc<B>(s, s->d_turn_active0, s->h_turn_active0.size())));
            RC(B::h2d(c, s->d_turn_active0, s->h_turn_active0.data(), s->h_turn_active0.size()));
        }
        RC(B::sync(c));  // (ks goes out of scope)
        s->init_images = true;
        RawVector<ActDyn>().swap(s->h_adyn);  // the device copy is the master from here on
    }
    // LinkHot{0,0,0,0,-1.0,0.0}: zeros, then blocked_until = -1 in every record (stride of 4 doubles)
    RC(B::memset(c, ar + d.off_hot, 0, sizeof(LinkHot) * L));
    RC(B::fill64(c, ar + d.off_hot + offsetof(LinkHot, blocked_until), 0xBFF0000000000000ull /* -1.0 */, L, sizeof(LinkHot) / 8));
    RC(B::memset(c, d.lacc, 0, sizeof(LinkAcc) * L));
    RC(B::d2d(c, d.avgtimes, d.link_hist, sizeof(double) * L * (size_t)s->n_periods));  // new LinkTime(*histtimes), B/link.h:122
    RC(B::fill64(c, d.arrival, 0xBFF0000000000000ull /* -1.0 */, T, 1));
    RC(B::memset(c, d.exit_log, 0xFF, sizeof(double2) * (size_t)std::max<long long>(1, s->total_log)));  // NaN = no exit yet
    RC(B::memset(c, d.turn_blocked, 0, (size_t)std::max(1, d.n_turnings)));
    if (d.turn_active) {
        RC(B::d2d(c, d.turn_active, s->d_turn_active0, s->h_turn_active0.size()));
        RC(B::memset(c, d.sig_next, 0, sizeof(int) * (size_t)std::max(1, s->n_sig_controls)));
        RC(B::memset(c, d.sig_endsub, 0, sizeof(int) * (size_t)(s->n_sig_events + 1)));
    }
    RC(B::memset(c, d.trip_parked, 0, T));
    RC(B::memset(c, d.od_next, 0, sizeof(int) * (size_t)std::max(1, s->n_odpairs)));
    // virtual-queue heads are absolute positions in origin_trips: each origin starts at its own first trip
    if (s->n_origins) RC(B::d2d(c, d.origin_vq_head, d.origin_trip_off, sizeof(int) * (size_t)s->n_origins));
    RC(B::fill64(c, d.srv_seed, (unsigned long long)s->randseed, (size_t)std::max(1, s->n_servers_total), 1));
    if (A) RC(B::d2d(c, d.adyn, s->d_adyn0, sizeof(ActDyn) * A));
    RC(B::d2d(c, ar + d.off_kt, s->d_key_t0, sizeof(double) * N));
    RC(B::d2d(c, ar + d.off_kslot[0], s->d_kslot0, sizeof(KeySlot) * N));
    RC(B::memset(c, ar + d.off_kslot[1], 0, sizeof(KeySlot) * N));
    RC(B::memset(c, ar + d.off_kver, 0, sizeof(unsigned) * N));
    RC(B::memset(c, ar + d.off_last, 0, sizeof(NodeLast) * N));  // t = 0: no execution instant yet (all times are >= 0.1)
    RC(B::memset(c, ar + s->off_sync, 0, sizeof(unsigned long long) * 64));
    RC(B::memset(c, d.counters, 0, sizeof(unsigned long long) * 64));
    RC(B::memset(c, d.mailbox, 0, sizeof(unsigned) * s->mailbox_words));
    if (d.trip_stop_off) {
        RC(B::memset(c, d.bus_next, 0, sizeof(int) * (size_t)std::max(1, s->n_stops)));
        RC(B::memset(c, d.slink_cur, 0, sizeof(int) * (size_t)std::max(1, s->n_slinks)));
        RC(B::memset(c, s->d_dep_off, 0, sizeof(int) * ((size_t)s->n_slinks + 1)));
        s->pending.clear();
        s->h_bus_slink.assign((size_t)std::max(1, s->n_stops), -1);
        s->dep_seq = s->visits_read = 0;
    }
    s->ran_until = 0.0;
    s->accum_ms = 0.0;
    RC(B::sync(c));
#undef RC
    s->ran = false;
    s->stats = mz_sim_stats{};
    return MZ_OK;
}


// Network::step: node visits until no node of this rank has an event inside the horizon left, then the single event
// beyond it (H8). B::run_visits does the former — the CUDA backend inside one persistent kernel; with several ranks every
// rank runs its own kernel and the nodes synchronise pairwise through the published keys in the arenas.
template <class B>
int sim_run(SimHost<B> *s)
{
    MZ_REQUIRE(s, MZ_EINVAL, "null handle");
    MZ_REQUIRE(!s->ran, MZ_ESTATE, "mz_sim_run already executed; call mz_sim_reset first");
    MZ_REQUIRE(s->n_ranks == 1 || s->peers_attached, MZ_ESTATE, "multi-rank engine: attach the peer arenas first");
    if (int rc = B::set_device(s->device)) return rc;
    auto w0 = std::chrono::steady_clock::now();
    const SimDev &d = s->dev;
    long long steps = 0, launches = 0;
    if (int rc = B::timer_start(s->ctx)) return rc;
    if (int rc = B::run_visits(s, &steps, &launches)) return rc;
    // "while (time < runtime) time = next()": the first event at/after the horizon still executes.
    // Every rank holds (or can read) all keys of its own nodes; the global minimum is found by the caller for several
    // ranks (sim_final_key / sim_final_event), here directly.
    double end_time = 0.0;
    if (s->n_ranks == 1) {
        Key best;
        if (int rc = sim_min_key(s, &best)) return rc;
        const bool have = best.id != 0x7fffffff && std::isfinite(best.t);
        // a MatrixAction or the poll of an inactive ODaction may be the event that ends the loop (MzNet::phantom_final_t)
        if (s->phantom_t > 0.0 && (!have || s->phantom_t < best.t || (s->phantom_t == best.t && s->phantom_tp < best.tp))) {
            end_time = s->phantom_t;
        } else if (have) {
            if (int rc = B::launch_final(s->ctx, d, best.id)) return rc;
            ++launches;
            end_time = best.t;
        }
    }
    double ms = 0.0;
    if (int rc = B::timer_stop(s->ctx, &ms)) return rc;
    ms += s->accum_ms;  // earlier slices of this run (mz_sim_run_until)
    unsigned long long c[64];
    if (int rc = B::d2h_sync(s->ctx, c, d.counters, sizeof(c))) return rc;
    for (int i = 0; i < 56; ++i) s->prof_ev[i] = c[8 + i];
    s->stats.n_stop_visits = (int64_t)c[6];
    s->stats.n_bus_refused = (int64_t)c[7];
    s->stats.n_traversals = (int64_t)c[0];
    s->stats.n_exits = (int64_t)c[1];
    s->stats.n_events = (int64_t)c[2] + (s->rank == 0 ? s->n_init_events : 0);
    s->stats.n_arrived = (int64_t)c[3];
    s->stats.n_supersteps = (int64_t)c[4];  // node visits that executed events (there are no global steps any more)
    (void)steps;
    s->stats.n_launches = launches;
    s->stats.kernel_ms = ms;
    s->stats.end_time = end_time;
    s->stats.total_ms = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - w0).count();
    s->ran = true;
    return MZ_OK;
}

// ---- transit overlay hooks -------------------------------------------------------------------
// One slice of the event loop: every event with time < min(t, runtime). The engine's state simply stays where the
// persistent kernel left it; the next slice (or sim_run) carries on.
template <class B>
int sim_run_until(SimHost<B> *s, double t)
{
    MZ_REQUIRE(s, MZ_EINVAL, "null handle");
    MZ_REQUIRE(!s->ran, MZ_ESTATE, "the run is complete; call mz_sim_reset first");
    MZ_REQUIRE(s->n_ranks == 1, MZ_ELIMIT, "mz_sim_run_until: single-GPU engines only");
    MZ_REQUIRE(t >= s->ran_until, MZ_EINVAL, "mz_sim_run_until(%g): the engine has already run to %g", t, s->ran_until);
    if (int rc = B::set_device(s->device)) return rc;
    const double until = std::min(t, s->horizon);
    long long steps = 0, launches = 0;
    if (int rc = B::timer_start(s->ctx)) return rc;
    s->dev.runtime = until;
    const int rc_run = B::run_visits(s, &steps, &launches);
    s->dev.runtime = s->horizon;
    if (rc_run) return rc_run;
    double ms = 0.0;
    if (int rc = B::timer_stop(s->ctx, &ms)) return rc;
    s->accum_ms += ms;
    s->stats.n_launches += launches;
    s->ran_until = until;
    return MZ_OK;
}

// the stop visits booked since the last call
template <class B>
int sim_stop_arrivals(SimHost<B> *s, mz_stop_visit *out, int64_t cap, int64_t *n_out)
{
    MZ_REQUIRE(s && n_out, MZ_EINVAL, "null argument");
    *n_out = 0;
    if (!s->dev.trip_stop_off) return MZ_OK;
    if (int rc = B::set_device(s->device)) return rc;
    unsigned long long cnt = 0;
    if (int rc = B::d2h_sync(s->ctx, &cnt, s->dev.counters + 6, sizeof(cnt))) return rc;
    const long long avail = (long long)cnt - s->visits_read;
    *n_out = avail;
    if (!out || avail == 0) return MZ_OK;
    MZ_REQUIRE(cap >= avail, MZ_ELIMIT, "%lld stop visits wait, buffer holds %lld", avail, (long long)cap);
    static_assert(sizeof(mz_stop_visit) == sizeof(StopVisit), "mz_stop_visit layout");
    if (int rc = B::d2h_sync(s->ctx, out, s->dev.stop_out + s->visits_read, sizeof(StopVisit) * (size_t)avail)) return rc;
    for (long long i = 0; i < avail; ++i) {
        const int s0 = s->h_trip_stop_off[out[i].veh];
        const int sl = (int)(std::lower_bound(s->h_slink_link.begin(), s->h_slink_link.end(), out[i].link) - s->h_slink_link.begin());
        s->h_bus_slink[(size_t)s0] = sl;  // where the bus waits for its departure
    }
    s->visits_read += avail;
    return MZ_OK;
}

// Departures booked by the host's transit model. The pending lists of all stop links are rewritten (what the engine has
// consumed — everything before the time it has run to — is dropped), each touched ACT_STOP action gets the time of its
// link's first pending departure, and a node whose published key lies beyond it is re-published with that time.
template <class B>
int sim_stop_departures(SimHost<B> *s, int64_t n, const int32_t *veh, const double *t_depart)
{
    MZ_REQUIRE(s && (n == 0 || (veh && t_depart)), MZ_EINVAL, "null argument");
    MZ_REQUIRE(s->dev.trip_stop_off, MZ_ESTATE, "the network has no stops");
    MZ_REQUIRE(!s->ran, MZ_ESTATE, "the run is complete");
    if (int rc = B::set_device(s->device)) return rc;
    typedef typename SimHost<B>::Dep Dep;
    for (int64_t i = 0; i < n; ++i) {
        MZ_REQUIRE(veh[i] >= 0 && veh[i] < s->n_trips, MZ_EINVAL, "departure %lld: bad vehicle", (long long)i);
        const int s0 = s->h_trip_stop_off[veh[i]];
        MZ_REQUIRE(s0 < s->h_trip_stop_off[veh[i] + 1] && s->h_bus_slink[(size_t)s0] >= 0, MZ_ESTATE,
                   "departure %lld: vehicle %d is not at a stop (drain mz_sim_stop_arrivals first)", (long long)i, veh[i]);
        MZ_REQUIRE(t_depart[i] >= s->ran_until, MZ_EINVAL, "departure %lld at %.9g precedes the time the engine has run to (%.9g)",
                   (long long)i, t_depart[i], s->ran_until);
        s->pending.push_back(Dep{t_depart[i], s->dep_seq++, veh[i], s->h_bus_slink[(size_t)s0]});
        s->h_bus_slink[(size_t)s0] = -1;
    }
    // consumed: executed in an earlier slice (events run while time < ran_until)
    s->pending.erase(std::remove_if(s->pending.begin(), s->pending.end(), [&](const Dep &d) { return d.t < s->ran_until; }),
                     s->pending.end());
    std::sort(s->pending.begin(), s->pending.end(), [](const Dep &a, const Dep &b) {
        if (a.slink != b.slink) return a.slink < b.slink;
        if (a.t != b.t) return a.t < b.t;
        return a.seq < b.seq;
    });
    MZ_REQUIRE((long long)s->pending.size() <= (long long)s->n_stops, MZ_ESTATE, "more pending departures than stop visits");
    std::vector<int> off((size_t)s->n_slinks + 1, 0), dv(s->pending.size());
    std::vector<double> dt(s->pending.size());
    for (size_t i = 0; i < s->pending.size(); ++i) {
        off[(size_t)s->pending[i].slink + 1]++;
        dv[i] = s->pending[i].veh;
        dt[i] = s->pending[i].t;
    }
    for (int i = 0; i < s->n_slinks; ++i) off[(size_t)i + 1] += off[(size_t)i];
    typename B::Ctx &c = s->ctx;
    const SimDev &d = s->dev;
    if (int rc = B::h2d(c, s->d_dep_off, off.data(), sizeof(int) * off.size())) return rc;
    if (!dv.empty()) {
        if (int rc = B::h2d(c, s->d_dep_veh, dv.data(), sizeof(int) * dv.size())) return rc;
        if (int rc = B::h2d(c, s->d_dep_t, dt.data(), sizeof(double) * dt.size())) return rc;
    }
    if (int rc = B::memset(c, d.slink_cur, 0, sizeof(int) * (size_t)std::max(1, s->n_slinks))) return rc;
    if (int rc = B::sync(c)) return rc;
    char *ar = (char *)s->arena;
    for (int sl = 0; sl < s->n_slinks; ++sl) {
        const bool any = off[(size_t)sl] < off[(size_t)sl + 1];
        const double t = any ? dt[(size_t)off[(size_t)sl]] : INFINITY;
        const int a = s->h_slink_act[(size_t)sl], node = s->h_act_node[(size_t)a];
        ActDyn ad{t, 0.0, 0, 0, 0.0};  // (t, parent time 0): among events of exactly equal time a departure goes first
        if (int rc = B::h2d(c, d.adyn + a, &ad, sizeof(ad))) return rc;
        if (!any) continue;
        double kt = 0.0;
        if (int rc = B::d2h_sync(c, &kt, ar + d.off_kt + sizeof(double) * (size_t)node, sizeof(kt))) return rc;
        if (!(t < kt)) continue;
        // publication ver+1 of the node's key, written the way node_key_publish does: slot, then time word and counter
        unsigned ver = 0;
        if (int rc = B::d2h_sync(c, &ver, ar + d.off_kver + sizeof(unsigned) * (size_t)node, sizeof(ver))) return rc;
        const unsigned v = ver + 1u;
        KeySlot ks{t, 0.0, 0, 0, 0.0};
        NodeLast nl;
        if (int rc = B::d2h_sync(c, &nl, ar + d.off_last + sizeof(NodeLast) * (size_t)node, sizeof(nl))) return rc;
        nl.ver = v;
        if (int rc = B::h2d(c, ar + d.off_kslot[v & 1u] + sizeof(KeySlot) * (size_t)node, &ks, sizeof(ks))) return rc;
        if (int rc = B::h2d(c, ar + d.off_kt + sizeof(double) * (size_t)node, &t, sizeof(t))) return rc;
        if (int rc = B::h2d(c, ar + d.off_kver + sizeof(unsigned) * (size_t)node, &v, sizeof(v))) return rc;
        if (int rc = B::h2d(c, ar + d.off_last + sizeof(NodeLast) * (size_t)node, &nl, sizeof(nl))) return rc;
    }
    return B::sync(c);
}

// smallest pending key among this rank's nodes (id = node, 0x7fffffff if none)
template <class B>
int sim_min_key(SimHost<B> *s, Key *out)
{
    const SimDev &d = s->dev;
    const size_t N = (size_t)d.n_nodes;
    std::vector<KeySlot> slot[2] = {std::vector<KeySlot>(N), std::vector<KeySlot>(N)};
    std::vector<unsigned> ver(N);
    char *ar = (char *)s->arena;
    if (int rc = B::d2h_sync(s->ctx, ver.data(), ar + d.off_kver, sizeof(unsigned) * N)) return rc;
    for (int k = 0; k < 2; ++k)
        if (int rc = B::d2h_sync(s->ctx, slot[k].data(), ar + d.off_kslot[k], sizeof(KeySlot) * N)) return rc;
    Key best{INFINITY, INFINITY, 0x7fffffff, 0x7fffffff};
    for (int i = 0; i < d.n_nodes; ++i) {
        if (s->h_node_home[i] != s->rank) continue;
        const KeySlot &ks = slot[ver[i] & 1u][i];
        Key k{ks.t, ks.tp, ks.sub, i};
        if (key_less(k, best)) best = k;
    }
    *out = best;
    return MZ_OK;
}

// Multi-rank epilogue: the caller all-gathers every rank's sim_min_key, picks the global minimum and tells the
// owning rank to execute it (H8); the other ranks pass final_node = -1 and only record the end time.
template <class B>
int sim_final_event(SimHost<B> *s, int final_node, double end_time)
{
    MZ_REQUIRE(s && s->ran, MZ_ESTATE, "mz_sim_run first");
    if (int rc = B::set_device(s->device)) return rc;
    if (final_node >= 0 && s->h_node_home[final_node] == s->rank) {
        if (int rc = B::launch_final(s->ctx, s->dev, final_node)) return rc;
        if (int rc = B::sync(s->ctx)) return rc;
        s->stats.n_launches++;
        unsigned long long c[8];
        if (int rc = B::d2h_sync(s->ctx, c, s->dev.counters, sizeof(c))) return rc;
        s->stats.n_traversals = (int64_t)c[0];
        s->stats.n_exits = (int64_t)c[1];
        s->stats.n_events = (int64_t)c[2] + (s->rank == 0 ? s->n_init_events : 0);
        s->stats.n_arrived = (int64_t)c[3];
    }
    s->stats.end_time = end_time;
    return MZ_OK;
}

template <class B>
int sim_get_exit_log(SimHost<B> *s, MzExit *out, int64_t cap, int64_t *n_out)
{
    MZ_REQUIRE(s && n_out, MZ_EINVAL, "null argument");
    if (int rc = B::set_device(s->device)) return rc;
    std::vector<double2> log((size_t)std::max<long long>(1, s->total_log));
    if (s->total_log)
        if (int rc = B::d2h_sync(s->ctx, log.data(), s->dev.exit_log, sizeof(double2) * (size_t)s->total_log)) return rc;
    int64_t cnt = 0;
    for (int v = 0; v < s->n_trips; ++v) {
        const int r = s->h_trip_route[v];
        const long long b = s->h_trip_log_off[v], e = s->h_trip_log_off[v + 1];
        for (long long i = b; i < e; ++i) {
            // not exited here: with one rank no later route position can have an exit either; a multi-rank engine
            // holds only the exits its own nodes performed, so it keeps scanning
            if (std::isnan(log[i].y)) {
                if (s->n_ranks == 1) break;
                continue;
            }
            if (out && cnt < cap)
                out[cnt] = MzExit{v + 1, s->h_route_links[s->h_route_off[r] + (int)(i - b)], log[i].x, log[i].y};
            ++cnt;
        }
    }
    *n_out = cnt;
    MZ_REQUIRE(!out || cnt <= cap, MZ_ELIMIT, "exit log holds %lld records, buffer %lld", (long long)cnt, (long long)cap);
    return MZ_OK;
}

// Link::end_of_simulation (B/link.cpp:118-141) on a host copy of Link::avgtimes: the running mean of the period that is
// still open is stored (historical time if nothing passed), periods left at 0 fall back to the historical time.
template <class B>
int sim_get_linktimes_final(SimHost<B> *s, double *avg)
{
    MZ_REQUIRE(s && avg, MZ_EINVAL, "null argument");
    if (int rc = B::set_device(s->device)) return rc;
    const size_t L = (size_t)s->n_links, P = (size_t)s->n_periods;
    if (L * P == 0) return MZ_OK;
    if (int rc = B::d2h_sync(s->ctx, avg, s->dev.avgtimes, sizeof(double) * L * P)) return rc;
    std::vector<LinkAcc> acc(L);
    if (int rc = B::d2h_sync(s->ctx, acc.data(), s->dev.lacc, sizeof(LinkAcc) * L)) return rc;
    const double *hist = s->h_link_hist.data();
    for (size_t l = 0; l < L; ++l) {
        // several ranks: a link's averages are kept by the rank of its downstream node; the other ranks report zeros, so
        // the caller's sum over ranks is the network's table
        if (s->n_ranks > 1 && s->h_node_home[(size_t)s->h_link_out[l]] != s->rank) {
            for (size_t i = 0; i < P; ++i) avg[l * P + i] = 0.0;
            continue;
        }
        const size_t cp = (size_t)acc[l].curr_period;
        if (cp < P) avg[l * P + cp] = acc[l].tmp_avg == 0.0 ? hist[l * P + cp] : acc[l].tmp_avg;
        for (size_t i = 0; i < P; ++i)
            if (avg[l * P + i] == 0.0) avg[l * P + i] = hist[l * P + i];
    }
    return MZ_OK;
}

template <class B>
int sim_create(int device, const MzNet *net, const SimPart &part, SimHost<B> **out)
{
    MZ_REQUIRE(out, MZ_EINVAL, "out handle pointer is null");
    *out = nullptr;
    {
        const bool timing = std::getenv("MZ_TIMING") != nullptr;
        auto t0 = std::chrono::steady_clock::now();
        if (int rc = sim_validate(net)) return rc;
        if (timing)
            std::fprintf(stderr, "[mz timing] %-28s %8.2f ms\n", "validate",
                         std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count());
    }
    MZ_REQUIRE(part.n_ranks >= 1 && part.n_ranks <= MZ_MAXR && part.rank >= 0 && part.rank < part.n_ranks, MZ_EINVAL,
               "bad partition: rank %d of %d (at most %d ranks)", part.rank, part.n_ranks, MZ_MAXR);
    MZ_REQUIRE(part.n_ranks == 1 || part.node_rank, MZ_EINVAL, "multi-rank engine needs node_rank[]");
    MZ_REQUIRE(part.n_ranks == 1 || !net->trip_stop_off, MZ_ELIMIT, "the transit overlay hooks (stops) need a single-GPU engine");
    if (part.n_ranks > 1)
        for (int i = 0; i < net->n_nodes; ++i)
            MZ_REQUIRE(part.node_rank[i] >= 0 && part.node_rank[i] < part.n_ranks, MZ_EINVAL, "node %d: bad rank", i);
    if (int rc = B::check_device(device)) return rc;
    SimHost<B> *s = new (std::nothrow) SimHost<B>;
    MZ_REQUIRE(s, MZ_ENOMEM, "out of host memory");
    s->device = device;
    s->n_ranks = part.n_ranks;
    s->rank = part.rank;
    if (part.n_ranks > 1) s->h_node_rank.assign(part.node_rank, part.node_rank + net->n_nodes);
    int rc = B::ctx_create(s->ctx);
    if (rc == MZ_OK) rc = sim_build(s, net);
    if (rc != MZ_OK) {
        sim_destroy(s);
        return rc;
    }
    *out = s;
    return MZ_OK;
}

template <class B>
int sim_destroy(SimHost<B> *s)
{
    if (!s) return MZ_OK;
    const auto t0 = std::chrono::steady_clock::now();
    B::set_device(s->device);
    B::detach_peers(s);
    for (void *p : s->allocs) B::free(s->ctx, p);
    if (s->arena) B::arena_free(s->ctx, s->arena, s->arena_bytes);
    B::ctx_destroy(s->ctx);
    delete s;
    if (std::getenv("MZ_TIMING"))
        std::fprintf(stderr, "[mz timing] %-28s %8.2f ms\n", "destroy",
                     std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count());
    return MZ_OK;
}

}  // namespace mzsim
