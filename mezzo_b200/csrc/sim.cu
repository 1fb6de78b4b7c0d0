// Hot path 1: the link update — Network::step's event loop recast as a conservative, window-synchronous step.
//
// Reference (B/ = /root/reference/branches/busmezzo/mezzo_lib/src/):
//   Network::init / step            B/network.cpp:7657-7754, 7792-7821      Eventlist           B/eventlist.h:58-71
//   TurnAction / Turning            B/turning.cpp:45-146, 237-265           Link                B/link.cpp:155-183, 320-390, 418-655, 670-687
//   Q                               B/q.cpp:52-129, 156-249, 346-361         Sdfunc              B/sdfunc.cpp:12-53
//   Server / Random                 B/server.cpp:32-50, B/Random.cpp:85-132  Daction / Origin    B/node.cpp:126-250, 637-673
//
// Execution model: every node (junction, origin, destination) is a logical process that owns
// its actions (TurnActions of its turnings, Dactions of its in-links, or the trips it generates). An action never wakes
// another action: each one re-books only itself. Therefore the earliest pending event key of a node is a hard lower
// bound on everything that node will ever do, and a node may safely execute all of its events that precede the
// earliest key among its neighbours (nodes sharing a link with it). One super-step = one window: all nodes whose own
// key is a local minimum advance up to their neighbours' keys; the windows are per-node and adaptive instead of a
// single global min-free-flow-time window because spill-back (Link::full), shock-wave unblocking and the running-segment
// density act with ZERO lookahead. Adjacent nodes are never active in the same super-step, so link
// queues (shared only by a link's two end nodes) need no atomics and the result equals the sequential event order.
// Ties in time are ordered like the reference's FIFO multimap by (time, time of the booking event, action index).
//
// Data layout in HBM (structure of arrays, no per-vehicle heap objects):
//   link  : length,lanes,sdfunc,maxcap,freeflow | q_off,q_head,q_size | blocked_until,nr_exits_blocked | period stats
//   queue : one ring buffer per link inside a single slab, entries {exit_time f64, veh i32, next_link i32} in LIST order
//   veh   : route,pos,length | entry,exit (current link) | meters | arrival   + exit log [Σ route lengths] {entry,exit}
//   action: {t, t_prev, kind, idx, aux} grouped by node (CSR)   node: neighbour CSR + published key (double-buffered)
#include <algorithm>
#include <chrono>
#include <cmath>
#include <map>
#include <set>

#include "common.h"

namespace {

struct QEnt {
    double exit;
    int veh;
    int next;  // Vehicle::nextlink() cached at insertion (route successor of the link the entry sits in)
};
static_assert(sizeof(QEnt) == 16, "queue entry must be 16 bytes");

struct Key {
    double t, tp;
    int id;
};
__host__ __device__ inline bool key_less(const Key &a, const Key &b)
{
    if (a.t != b.t) return a.t < b.t;
    if (a.tp != b.tp) return a.tp < b.tp;
    return a.id < b.id;
}

enum { ACT_TURN = 0, ACT_DEST = 1 };
enum { NODE_JUNCTION = 0, NODE_ORIGIN = 1, NODE_DEST = 2 };

// everything the kernels need, passed by value
struct SimDev {
    int n_nodes, n_links, n_turnings, n_trips, n_periods, n_acts;
    double runtime, periodlength;
    // static network
    const int *link_length, *link_lanes, *link_sdfunc, *link_maxcap, *link_qoff, *link_qcap;
    const double *link_freeflow, *link_hist;
    const int *sd_type;
    const double *sd_vfree, *sd_vmin, *sd_romax, *sd_romin, *sd_alpha, *sd_beta;
    const int *srv_type;
    const double *srv_mu, *srv_sd, *srv_delay;
    const int *turn_server, *turn_in, *turn_out, *turn_lookback;
    const int *route_off, *route_links;
    // node structure
    const int *node_type, *node_act_off, *node_nb_off, *node_nb, *node_origin;  // node_origin: origin index or -1
    const int *origin_trip_off, *origin_trips;
    // mutable link state
    int *q_head, *q_size, *nr_exits_blocked, *nr_passed, *tmp_passed, *curr_period;
    double *blocked_until, *avg_time, *tmp_avg, *avgtimes;
    QEnt *slab;
    // vehicles
    const int *trip_route;
    const double *trip_time, *trip_length;
    const long long *trip_log_off;
    int *veh_pos, *veh_meters;
    double *veh_entry, *veh_exit, *arrival;
    double2 *exit_log;  // {entry, exit} per (vehicle, route position)
    // actions
    double *act_t, *act_tp;
    const int *act_kind, *act_idx, *act_aux;
    unsigned char *turn_blocked;
    long long *srv_seed;
    // origins
    int *origin_next, *origin_vq_head;
    unsigned char *trip_parked;
    // published node keys (double buffered)
    double *key_t[2], *key_tp[2];
    // counters: [0] traversals [1] exits [2] events [3] arrived [4] alive flag [5] error flag
    unsigned long long *counters;
};

// ------------------------------------------------------------------------------------------------ Random / Server
__device__ __forceinline__ double urandom(long long &seed)  // B/Random.cpp:85-98
{
    const long long M = 2147483647LL, A = 48271LL, Q = M / A, R = M % A;
    long long s = A * (seed % Q) - R * (seed / Q);
    s = (s > 0) ? s : (s + M);
    seed = s;
    return (double)s / (double)M;
}
__device__ __forceinline__ double nrandom(long long &seed, double mean, double sd)  // B/Random.cpp:126-132
{
    const double r1 = urandom(seed), r2 = urandom(seed);
    const double r = -2.0 * log(r1);
    if (r > 0.0) return mean + sd * sqrt(r) * sin((3.14159265358979323846 * 2.0) * r2);
    return mean;
}
__device__ __forceinline__ double server_next(int type, double mu, double sd, double delay, long long &seed, double t)
{
    if (type == 0) return t;               // ConstServer   B/server.h:74
    if (type == 2) return t + mu + delay;  // DetServer     B/server.h:98
    const double x = nrandom(seed, mu, sd);  // Server::next  B/server.cpp:46-48
    return t + (x > 0.1 ? x : 0.1);
}
__device__ __forceinline__ double sd_speed(const SimDev &d, int f, double ro)  // B/sdfunc.cpp:12-19, 45-53
{
    if (ro <= d.sd_romin[f]) return d.sd_vfree[f];
    if (ro > d.sd_romax[f]) return d.sd_vmin[f];
    const double x = (ro - d.sd_romin[f]) / (d.sd_romax[f] - d.sd_romin[f]);
    if (d.sd_type[f] == 0) return d.sd_vmin[f] + (d.sd_vfree[f] - d.sd_vmin[f]) * (1 - x);
    return d.sd_vmin[f] + (d.sd_vfree[f] - d.sd_vmin[f]) * pow((1 - pow(x, d.sd_alpha[f])), d.sd_beta[f]);
}

// ------------------------------------------------------------------------------------------------ queue (ring buffer in list order)
struct QRef {
    QEnt *base;
    int cap, head, size;
};
__device__ __forceinline__ QRef q_open(const SimDev &d, int l)
{
    QRef q;
    q.base = d.slab + d.link_qoff[l];
    q.cap = d.link_qcap[l];
    q.head = d.q_head[l];
    q.size = d.q_size[l];
    return q;
}
__device__ __forceinline__ void q_close(const SimDev &d, int l, const QRef &q)
{
    d.q_head[l] = q.head;
    d.q_size[l] = q.size;
}
__device__ __forceinline__ QEnt &q_at(const QRef &q, int i)
{
    int p = q.head + i;
    if (p >= q.cap) p -= q.cap;
    return q.base[p];
}
// erase list position k: the k entries in front of it move one slot towards the tail, the head advances
__device__ __forceinline__ void q_erase(QRef &q, int k)
{
    for (int i = k; i > 0; --i) q_at(q, i) = q_at(q, i - 1);
    q.head = q.head + 1 == q.cap ? 0 : q.head + 1;
    q.size--;
}
// insert at list position p (0..size)
__device__ __forceinline__ void q_insert(QRef &q, int p, const QEnt &e)
{
    if (p >= q.size - p) {  // closer to the tail: shift the tail part back by one
        for (int i = q.size; i > p; --i) q_at(q, i) = q_at(q, i - 1);
        q.size++;
    } else {  // closer to the head: move the head one slot forward and shift the front part
        q.head = q.head == 0 ? q.cap - 1 : q.head - 1;
        q.size++;
        for (int i = 0; i < p; ++i) q_at(q, i) = q_at(q, i + 1);
    }
    q_at(q, p) = e;
}

__device__ __forceinline__ int veh_next_after(const SimDev &d, int v, int pos)
{
    const int r = d.trip_route[v];
    const int i = d.route_off[r] + pos + 1;
    return i < d.route_off[r + 1] ? d.route_links[i] : -1;
}

// ------------------------------------------------------------------------------------------------ Link
__device__ __forceinline__ bool link_full_t(const SimDev &d, int l, double t)  // Link::full(time) B/link.cpp:155-183
{
    const double bu = d.blocked_until[l];
    if (bu == -2.0) return true;
    if (d.q_size[l] >= d.link_maxcap[l]) {
        if (d.nr_exits_blocked[l] > 0) d.blocked_until[l] = -2.0;
        return true;
    }
    if (bu == -1.0) return false;
    if (bu <= t) {
        d.blocked_until[l] = -1.0;
        return false;
    }
    return true;
}

// Link::enter_veh (B/link.cpp:418-523) = density_running_only (670-687) + Sdfunc::speed + Q::enter_veh (B/q.cpp:52-72)
__device__ void link_enter_veh(const SimDev &d, int l, int v, double t, unsigned long long &n_trav)
{
    QRef q = q_open(d, l);
    // nr_running: elements from the first one with exit_time > t to the end (B/q.h:67-68)
    int first = 0;
    while (first < q.size && !(q_at(q, first).exit > t)) ++first;
    const int nr = q.size - first;
    // calc_space: vehicle lengths from the back while exit_time <= t (B/q.cpp:346-361)
    double space = 0.0;
    for (int i = q.size - 1; i >= 0; --i) {
        const QEnt &e = q_at(q, i);
        if (e.exit > t) break;
        space += d.trip_length[e.veh];
    }
    const int lanes = d.link_lanes[l], len = d.link_length[l];
    const double queue_space = space / lanes;
    const double running_length = len - queue_space;
    double ro = 0.0;
    if (nr && running_length) ro = nr / (running_length * (lanes / 1000.0));
    const double speed = sd_speed(d, d.link_sdfunc[l], ro);
    const double exit_time = t + (len / speed);
    const int pos = d.veh_pos[v] + 1;
    d.veh_pos[v] = pos;
    d.veh_entry[v] = t;
    d.veh_exit[v] = exit_time;
    QEnt e;
    e.exit = exit_time;
    e.veh = v;
    e.next = veh_next_after(d, v, pos);
    if (q.size == 0) {
        q_insert(q, 0, e);
    } else {
        int i = q.size - 1;
        while (q_at(q, i).exit > exit_time && i != 0) --i;
        q_insert(q, i + 1, e);  // insert(++viter): second place even if the head itself is later
    }
    q_close(d, l, q);
    ++n_trav;
}

// bookkeeping of a successful Link::exit_veh (B/link.cpp:536-581 / 601-648) + exit log
__device__ void link_exit_bookkeeping(const SimDev &d, int l, int v, double t, unsigned long long &n_exits)
{
    const double entrytime = d.veh_entry[v], traveltime = t - entrytime;
    const int np = d.nr_passed[l];
    d.avg_time[l] = (np * d.avg_time[l] + traveltime) / (np + 1);
    const int cp = d.curr_period[l];
    if ((cp + 1) * d.periodlength < entrytime) {
        double ta = d.tmp_avg[l];
        if (ta == 0.0) ta = cp < d.n_periods ? d.link_hist[(size_t)l * d.n_periods + cp] : 0.0;
        if (cp < d.n_periods) d.avgtimes[(size_t)l * d.n_periods + cp] = ta;
        d.curr_period[l] = cp + 1;
        d.tmp_avg[l] = 0.0;
        d.tmp_passed[l] = 0;
    } else {
        const int tp = d.tmp_passed[l];
        d.tmp_avg[l] = (tp * d.tmp_avg[l] + traveltime) / (tp + 1);
        d.tmp_passed[l] = tp + 1;
    }
    d.nr_passed[l] = np + 1;
    d.veh_meters[v] += d.link_length[l];
    d.exit_log[d.trip_log_off[v] + d.veh_pos[v]] = make_double2(entrytime, t);
    ++n_exits;
}

// Link::update_exit_times + Q::update_exit_times (B/link.cpp:320-390, B/q.cpp:74-129)
__device__ void link_update_exit_times(const SimDev &d, int l, double t, int next, int lookback)
{
    const double kjam = 142.0;
    const int f = d.link_sdfunc[l];
    const double k_dn = d.q_size[next] / (d.link_lanes[next] * (d.link_length[next] / 1000.0));  // Link::density
    const double v_exit = sd_speed(d, f, kjam);
    double v_dn = sd_speed(d, d.link_sdfunc[next], k_dn);
    double v_sw;
    if (v_dn <= v_exit) v_dn = v_exit - 1.0;
    if (k_dn >= d.sd_romax[f]) v_sw = v_exit - 1.0;
    else v_sw = (k_dn * v_dn) / (kjam - k_dn);
    if (v_sw >= v_exit) v_sw = v_exit - 1.0;
    if (v_sw == 0.0) v_sw = v_exit;
    QRef q = q_open(d, l);
    int cnt = 0, i = 0;
    double t0 = t;
    while (i < q.size) {
        QEnt &e = q_at(q, i);
        if (e.exit > t0) break;
        const double vl = d.trip_length[e.veh];
        if (cnt < lookback) {
            cnt++;
            if (e.next == next) {
                const double newtime = t0 + vl / v_sw + vl / v_exit;
                e.exit = newtime;
                d.veh_exit[e.veh] = newtime;
                t0 = newtime;
            }
            i++;
        } else {
            cnt++;
            const double newtime = t0 + vl / v_sw + vl / v_exit;
            const int lanes = d.link_lanes[l];
            for (int k = 0; k < lanes && i < q.size; ++k) {
                QEnt &e2 = q_at(q, i);
                e2.exit = newtime;
                d.veh_exit[e2.veh] = newtime;
                i++;
            }
            t0 = newtime;
        }
    }
    if (d.blocked_until[l] == -2.0) d.blocked_until[l] = (d.link_length[l] / v_sw) + t;
}

struct Tally {
    unsigned long long trav = 0, exits = 0, events = 0, arrived = 0;
};

// TurnAction::execute + Turning::process_veh (B/turning.cpp:237-265, 45-146); returns the re-booking time
__device__ double turn_execute(const SimDev &d, int k, double t, Tally &ty)
{
    const int in = d.turn_in[k], out = d.turn_out[k], sv = d.turn_server[k];
    double nexttime = t;
    bool ok = false;
    if (link_full_t(d, out, t)) {
        if (!d.turn_blocked[k]) {
            d.turn_blocked[k] = 1;
            d.nr_exits_blocked[in]++;
        }
        nexttime = t + 1.0;
    } else {
        if (d.turn_blocked[k]) {
            d.turn_blocked[k] = 0;
            link_update_exit_times(d, in, t, out, d.turn_lookback[k]);
            d.nr_exits_blocked[in]--;
        }
        QRef q = q_open(d, in);
        if (q.size == 0) {
            nexttime = t + d.link_freeflow[in];
        } else {
            // Q::exit_veh(time, nextlink, lookback): first vehicle in LIST order heading for `out` (B/q.cpp:156-222)
            int n = 0;
            while (n < q.size && q_at(q, n).next != out) ++n;
            if (n == q.size) {
                nexttime = t + d.link_freeflow[in];  // next_action = time + freeflowtime
            } else {
                const QEnt e = q_at(q, n);
                if (e.exit <= t) {
                    if (n < d.turn_lookback[k]) {
                        q_erase(q, n);
                        q_close(d, in, q);
                        link_exit_bookkeeping(d, in, e.veh, t, ty.exits);
                        link_enter_veh(d, out, e.veh, t + d.srv_delay[sv], ty.trav);
                        ok = true;
                    } else {
                        nexttime = t + 1.0 * (n / d.link_lanes[out]);  // integer division, B/q.cpp:204
                    }
                } else {
                    nexttime = e.exit;
                }
            }
        }
    }
    if (ok) {
        long long seed = d.srv_seed[sv];
        const double nt = server_next(d.srv_type[sv], d.srv_mu[sv], d.srv_sd[sv], d.srv_delay[sv], seed, t);
        d.srv_seed[sv] = seed;
        return nt;
    }
    return nexttime < t ? t + 0.1 : nexttime;
}

// Daction::execute (B/node.cpp:637-673) with Q::exit_veh(time) (B/q.cpp:227-249)
__device__ double dest_execute(const SimDev &d, int dst_server_slot, int sv, int l, double t, Tally &ty)
{
    QRef q = q_open(d, l);
    if (q.size == 0) return t + d.link_freeflow[l];
    const QEnt e = q_at(q, 0);
    if (e.exit <= t) {
        q_erase(q, 0);
        q_close(d, l, q);
        link_exit_bookkeeping(d, l, e.veh, t, ty.exits);
        d.arrival[e.veh] = t;  // Vehicle::report(time)
        ++ty.arrived;
        long long seed = d.srv_seed[dst_server_slot];
        double nt;
        if (sv >= 0) nt = server_next(d.srv_type[sv], d.srv_mu[sv], d.srv_sd[sv], d.srv_delay[sv], seed, t);
        else nt = server_next(1, 0.1, 0.5, 0.001, seed, t);  // Destination's own Server(-20000,0,0.1,0.5,0.001)
        d.srv_seed[dst_server_slot] = seed;
        return nt;
    }
    return e.exit;
}

// Origin::insert_veh / transfer_veh with the InputLink virtual queue (B/node.cpp:126-250, B/link.cpp:949-995).
// The virtual queue of origin o is the set of its trips in [vq_head, next) whose `parked` flag is set, in trip order.
__device__ void origin_insert(const SimDev &d, int o, int slot, double t, Tally &ty)
{
    const int *trips = d.origin_trips;
    const int v = trips[slot];
    const int first = d.route_links[d.route_off[d.trip_route[v]]];
    int vqh = d.origin_vq_head[o];
    while (vqh < slot && !d.trip_parked[trips[vqh]]) ++vqh;  // drop leading non-parked trips
    const bool vq_empty = vqh == slot;
    if (d.q_size[first] >= d.link_maxcap[first]) {  // link->full(): park
        d.trip_parked[v] = 1;
        d.veh_entry[v] = t;
        d.veh_exit[v] = t;
    } else if (vq_empty) {
        link_enter_veh(d, first, v, t, ty.trav);
    } else {
        d.trip_parked[v] = 1;
        d.veh_entry[v] = t;
        d.veh_exit[v] = t;
        // transfer_veh: head-first, every parked vehicle whose first link is `first`, while the link is not full
        for (int i = vqh; i <= slot; ++i) {
            if (d.q_size[first] >= d.link_maxcap[first]) break;
            const int w = trips[i];
            if (!d.trip_parked[w]) continue;
            if (d.route_links[d.route_off[d.trip_route[w]]] != first) continue;
            d.trip_parked[w] = 0;
            link_enter_veh(d, first, w, t, ty.trav);
        }
    }
    d.origin_vq_head[o] = vqh;
}

// ------------------------------------------------------------------------------------------------ the super-step
__device__ __forceinline__ Key node_min_key(const SimDev &d, int n, int &which)
{
    Key best;
    best.t = INFINITY;
    best.tp = INFINITY;
    best.id = 0x7fffffff;
    which = -1;
    const int type = d.node_type[n];
    if (type == NODE_ORIGIN) {
        const int o = d.node_origin[n];
        const int nx = d.origin_next[o];
        if (d.origin_trip_off[o] + nx < d.origin_trip_off[o + 1]) {
            best.t = d.trip_time[d.origin_trips[d.origin_trip_off[o] + nx]];
            best.tp = 0.0;
            best.id = n;
            which = 0;
        }
        return best;
    }
    for (int a = d.node_act_off[n]; a < d.node_act_off[n + 1]; ++a) {
        Key k;
        k.t = d.act_t[a];
        k.tp = d.act_tp[a];
        k.id = a;
        if (key_less(k, best)) {
            best = k;
            which = a;
        }
    }
    return best;
}

// executes the earliest event of node n (identified by `which`); updates that action's key
__device__ void node_execute(const SimDev &d, int n, int which, double t, Tally &ty)
{
    ++ty.events;
    const int type = d.node_type[n];
    if (type == NODE_ORIGIN) {
        const int o = d.node_origin[n];
        const int nx = d.origin_next[o];
        origin_insert(d, o, d.origin_trip_off[o] + nx, t, ty);
        d.origin_next[o] = nx + 1;
        return;
    }
    double nt;
    if (d.act_kind[which] == ACT_TURN) nt = turn_execute(d, d.act_idx[which], t, ty);
    else nt = dest_execute(d, d.act_aux[which], d.act_aux[which + d.n_acts], d.act_idx[which], t, ty);
    d.act_tp[which] = t;
    d.act_t[which] = nt;
}

// One conservative window. `final_node` >= 0 selects the clean-up launch that executes exactly one event (H8).
__global__ void __launch_bounds__(128) k_sim_superstep(SimDev d, int parity, int final_node)
{
    const int n = blockIdx.x * blockDim.x + threadIdx.x;
    if (n >= d.n_nodes) return;
    const double *kt = d.key_t[parity], *ktp = d.key_tp[parity];
    double *ot = d.key_t[parity ^ 1], *otp = d.key_tp[parity ^ 1];
    Key mine;
    mine.t = kt[n];
    mine.tp = ktp[n];
    mine.id = n;
    Tally ty;
    if (final_node >= 0) {
        if (n == final_node) {
            int which;
            Key k = node_min_key(d, n, which);
            if (which >= 0) node_execute(d, n, which, k.t, ty);
        }
    } else if (mine.t < d.runtime) {
        // the window: everything before the earliest pending key of any neighbour
        Key bound;
        bound.t = INFINITY;
        bound.tp = INFINITY;
        bound.id = 0x7fffffff;
        for (int i = d.node_nb_off[n]; i < d.node_nb_off[n + 1]; ++i) {
            const int m = d.node_nb[i];
            Key k;
            k.t = kt[m];
            k.tp = ktp[m];
            k.id = m;
            if (key_less(k, bound)) bound = k;
        }
        if (key_less(mine, bound)) {
            for (;;) {
                int which;
                Key k = node_min_key(d, n, which);
                if (which < 0 || !(k.t < d.runtime)) break;
                Key kn = k;
                kn.id = n;  // keys are compared across nodes by node index
                if (!key_less(kn, bound)) break;
                node_execute(d, n, which, k.t, ty);
            }
        }
    }
    int which;
    const Key k = node_min_key(d, n, which);
    ot[n] = k.t;
    otp[n] = k.tp;
    if (k.t < d.runtime) d.counters[4] = 1;  // somebody still has work inside the horizon
    if (ty.events) {
        atomicAdd(d.counters + 2, ty.events);
        if (ty.trav) atomicAdd(d.counters + 0, ty.trav);
        if (ty.exits) atomicAdd(d.counters + 1, ty.exits);
        if (ty.arrived) atomicAdd(d.counters + 3, ty.arrived);
    }
}

}  // namespace

// ================================================================================================ host side
struct mz_sim {
    int device = 0;
    SimDev dev{};
    // owned device buffers
    std::vector<void *> allocs;
    // host copies needed for reset / init
    std::vector<double> h_act_t, h_act_tp, h_key_t, h_key_tp, h_link_hist;
    std::vector<int> h_act_node;
    int n_trips = 0, n_links = 0, n_periods = 0, n_nodes = 0, n_acts = 0, n_origins = 0, n_servers_total = 0;
    long long total_log = 0, slab_entries = 0;
    long long randseed = 0;
    long long n_init_events = 0;
    std::vector<long long> h_trip_log_off;
    cudaStream_t s_own = nullptr, stream = nullptr;
    cudaEvent_t ev0 = nullptr, ev1 = nullptr;
    mz_sim_stats stats{};
    bool ran = false;
    int check_every = 32;
};

namespace {

template <class T>
int dev_upload(mz_sim *s, const T *&dst, const std::vector<T> &src)
{
    T *p = nullptr;
    MZ_CUDA(cudaMalloc((void **)&p, std::max<size_t>(1, src.size()) * sizeof(T)));
    s->allocs.push_back(p);
    if (!src.empty()) MZ_CUDA(cudaMemcpy(p, src.data(), src.size() * sizeof(T), cudaMemcpyHostToDevice));
    dst = p;
    return MZ_OK;
}
template <class T>
int dev_alloc(mz_sim *s, T *&dst, size_t n)
{
    T *p = nullptr;
    MZ_CUDA(cudaMalloc((void **)&p, std::max<size_t>(1, n) * sizeof(T)));
    s->allocs.push_back(p);
    dst = p;
    return MZ_OK;
}
template <class T>
std::vector<T> vec(const T *p, size_t n)
{
    return std::vector<T>(p, p + n);
}

double host_sd_speed0(const MzNet *n, int f)
{
    // Sdfunc::speed(0.0): 0 <= romin always (reader asserts romin >= 0) ⇒ vfree
    return 0.0 <= n->sd_romin[f] ? n->sd_vfree[f] : n->sd_vmin[f];
}

int sim_validate(const MzNet *n)
{
    MZ_REQUIRE(n, MZ_EINVAL, "null network");
    MZ_REQUIRE(n->n_nodes > 0 && n->n_links > 0 && n->n_sdfuncs > 0 && n->n_periods > 0, MZ_EINVAL,
               "network needs nodes, links, sdfuncs and link-time periods");
    MZ_REQUIRE(n->randseed != 0, MZ_EINVAL, "randseed must be != 0 (0 means clock seeding in the reference)");
    MZ_REQUIRE(n->use_giveway == 0, MZ_ELIMIT, "use_giveway is not supported (SURVEY.md: off by default)");
    MZ_REQUIRE(n->server_type != 3 && n->server_type != 4, MZ_ELIMIT,
               "server_type 3/4 (lognormal / loglogistic servers) are not supported");
    MZ_REQUIRE(n->standard_veh_length > 0 && n->periodlength > 0 && n->runtime > 0, MZ_EINVAL, "bad parameters");
    for (int i = 0; i < n->n_links; ++i) {
        MZ_REQUIRE(n->link_length[i] > 0 && n->link_lanes[i] > 0, MZ_EINVAL, "link %d: length and lanes must be > 0", i);
        MZ_REQUIRE(n->link_sdfunc[i] >= 0 && n->link_sdfunc[i] < n->n_sdfuncs, MZ_EINVAL, "link %d: bad sdfunc", i);
        MZ_REQUIRE(n->link_in_node[i] >= 0 && n->link_in_node[i] < n->n_nodes && n->link_out_node[i] >= 0 &&
                       n->link_out_node[i] < n->n_nodes, MZ_EINVAL, "link %d: bad end node", i);
    }
    for (int i = 0; i < n->n_servers; ++i)
        MZ_REQUIRE(n->srv_type[i] >= 0 && n->srv_type[i] <= 2, MZ_ELIMIT, "server %d: type %d not supported (0,1,2 are)",
                   i, n->srv_type[i]);
    for (int k = 0; k < n->n_turnings; ++k) {
        MZ_REQUIRE(n->turn_in[k] >= 0 && n->turn_in[k] < n->n_links && n->turn_out[k] >= 0 && n->turn_out[k] < n->n_links,
                   MZ_EINVAL, "turning %d: bad link", k);
        MZ_REQUIRE(n->turn_server[k] >= 0 && n->turn_server[k] < n->n_servers, MZ_EINVAL, "turning %d: bad server", k);
        MZ_REQUIRE(n->link_out_node[n->turn_in[k]] == n->turn_node[k] && n->link_in_node[n->turn_out[k]] == n->turn_node[k],
                   MZ_EINVAL, "turning %d: its links do not meet at its node", k);
    }
    for (int r = 0; r < n->n_routes; ++r) {
        MZ_REQUIRE(n->route_off[r] < n->route_off[r + 1], MZ_EINVAL, "route %d is empty", r);
        std::set<int> seen;
        for (int i = n->route_off[r]; i < n->route_off[r + 1]; ++i) {
            MZ_REQUIRE(n->route_links[i] >= 0 && n->route_links[i] < n->n_links, MZ_EINVAL, "route %d: bad link", r);
            // Route::nextlink searches the FIRST occurrence of the current link (B/route.cpp:139-149)
            MZ_REQUIRE(seen.insert(n->route_links[i]).second, MZ_ELIMIT, "route %d visits link %d twice", r, n->route_links[i]);
        }
    }
    for (int v = 0; v < n->n_trips; ++v) {
        MZ_REQUIRE(n->trip_route[v] >= 0 && n->trip_route[v] < n->n_routes, MZ_EINVAL, "trip %d: bad route", v);
        MZ_REQUIRE(n->trip_origin[v] >= 0 && n->trip_origin[v] < n->n_origins, MZ_EINVAL, "trip %d: bad origin", v);
        MZ_REQUIRE(v == 0 || n->trip_time[v] >= n->trip_time[v - 1], MZ_EINVAL, "trips must be in time order");
        const int first = n->route_links[n->route_off[n->trip_route[v]]];
        MZ_REQUIRE(n->link_in_node[first] == n->origin_node[n->trip_origin[v]], MZ_EINVAL,
                   "trip %d: route does not start at its origin", v);
    }
    return MZ_OK;
}

// (Re)initialises all mutable device state: Network::init's first events and empty queues.
int sim_reset_state(mz_sim *s)
{
    SimDev &d = s->dev;
    MZ_CUDA(cudaSetDevice(s->device));
    cudaStream_t st = s->stream;
    MZ_CUDA(cudaMemsetAsync(d.q_head, 0, sizeof(int) * s->n_links, st));
    MZ_CUDA(cudaMemsetAsync(d.q_size, 0, sizeof(int) * s->n_links, st));
    MZ_CUDA(cudaMemsetAsync(d.nr_exits_blocked, 0, sizeof(int) * s->n_links, st));
    MZ_CUDA(cudaMemsetAsync(d.nr_passed, 0, sizeof(int) * s->n_links, st));
    MZ_CUDA(cudaMemsetAsync(d.tmp_passed, 0, sizeof(int) * s->n_links, st));
    MZ_CUDA(cudaMemsetAsync(d.curr_period, 0, sizeof(int) * s->n_links, st));
    MZ_CUDA(cudaMemsetAsync(d.avg_time, 0, sizeof(double) * s->n_links, st));
    MZ_CUDA(cudaMemsetAsync(d.tmp_avg, 0, sizeof(double) * s->n_links, st));
    std::vector<double> bu(s->n_links, -1.0);
    MZ_CUDA(cudaMemcpyAsync(d.blocked_until, bu.data(), sizeof(double) * s->n_links, cudaMemcpyHostToDevice, st));
    MZ_CUDA(cudaMemcpyAsync(d.avgtimes, s->h_link_hist.data(), sizeof(double) * s->h_link_hist.size(),
                            cudaMemcpyHostToDevice, st));  // avgtimes = new LinkTime(*histtimes), B/link.h:122
    MZ_CUDA(cudaMemsetAsync(d.veh_pos, 0xFF, sizeof(int) * std::max(1, s->n_trips), st));  // -1 = not entered
    MZ_CUDA(cudaMemsetAsync(d.veh_meters, 0, sizeof(int) * std::max(1, s->n_trips), st));
    std::vector<double> arr(std::max(1, s->n_trips), -1.0);
    MZ_CUDA(cudaMemcpyAsync(d.arrival, arr.data(), sizeof(double) * arr.size(), cudaMemcpyHostToDevice, st));
    MZ_CUDA(cudaMemsetAsync(d.exit_log, 0xFF, sizeof(double2) * std::max<long long>(1, s->total_log), st));  // NaN = no exit yet
    MZ_CUDA(cudaMemsetAsync(d.turn_blocked, 0, std::max(1, d.n_turnings), st));
    MZ_CUDA(cudaMemsetAsync(d.trip_parked, 0, std::max(1, s->n_trips), st));
    MZ_CUDA(cudaMemsetAsync(d.origin_next, 0, sizeof(int) * std::max(1, s->n_origins), st));
    MZ_CUDA(cudaMemsetAsync(d.origin_vq_head, 0, sizeof(int) * std::max(1, s->n_origins), st));
    std::vector<long long> seeds(std::max(1, s->n_servers_total), s->randseed);
    MZ_CUDA(cudaMemcpyAsync(d.srv_seed, seeds.data(), sizeof(long long) * seeds.size(), cudaMemcpyHostToDevice, st));
    MZ_CUDA(cudaMemcpyAsync(d.act_t, s->h_act_t.data(), sizeof(double) * std::max<size_t>(1, s->h_act_t.size()),
                            cudaMemcpyHostToDevice, st));
    MZ_CUDA(cudaMemcpyAsync(d.act_tp, s->h_act_tp.data(), sizeof(double) * std::max<size_t>(1, s->h_act_tp.size()),
                            cudaMemcpyHostToDevice, st));
    MZ_CUDA(cudaMemcpyAsync(d.key_t[0], s->h_key_t.data(), sizeof(double) * s->n_nodes, cudaMemcpyHostToDevice, st));
    MZ_CUDA(cudaMemcpyAsync(d.key_tp[0], s->h_key_tp.data(), sizeof(double) * s->n_nodes, cudaMemcpyHostToDevice, st));
    MZ_CUDA(cudaMemsetAsync(d.counters, 0, sizeof(unsigned long long) * 8, st));
    MZ_CUDA(cudaStreamSynchronize(st));
    s->ran = false;
    s->stats = mz_sim_stats{};
    return MZ_OK;
}

int sim_build(mz_sim *s, const MzNet *n)
{
    SimDev &d = s->dev;
    const int L = n->n_links, N = n->n_nodes, P = n->n_periods;
    s->n_links = L;
    s->n_nodes = N;
    s->n_periods = P;
    s->n_trips = n->n_trips;
    s->n_origins = n->n_origins;
    s->randseed = n->randseed;
    d.n_nodes = N;
    d.n_links = L;
    d.n_turnings = n->n_turnings;
    d.n_trips = n->n_trips;
    d.n_periods = P;
    d.runtime = n->runtime;
    d.periodlength = n->periodlength;

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

    // ---- node roles
    std::vector<int> node_type(N, NODE_JUNCTION), node_origin(N, -1);
    for (int o = 0; o < n->n_origins; ++o) {
        MZ_REQUIRE(n->origin_node[o] >= 0 && n->origin_node[o] < N, MZ_EINVAL, "origin %d: bad node", o);
        node_type[n->origin_node[o]] = NODE_ORIGIN;
        node_origin[n->origin_node[o]] = o;
    }
    std::vector<int> dest_of_node(N, -1);
    for (int k = 0; k < n->n_dests; ++k) {
        MZ_REQUIRE(n->dest_node[k] >= 0 && n->dest_node[k] < N, MZ_EINVAL, "destination %d: bad node", k);
        MZ_REQUIRE(node_type[n->dest_node[k]] == NODE_JUNCTION, MZ_EINVAL, "node is both origin and destination");
        node_type[n->dest_node[k]] = NODE_DEST;
        dest_of_node[n->dest_node[k]] = k;
    }

    // ---- actions in Network::init order with their first booking (B/network.cpp:7657-7731, B/turning.cpp:207-216)
    struct HAct { int node, kind, idx, aux, sv; double t, tp; };
    std::vector<HAct> acts;
    double initvalue = 0.1;
    for (int k = 0; k < n->n_turnings; ++k) {
        const int cnt = std::min(n->link_lanes[n->turn_in[k]], n->link_lanes[n->turn_out[k]]);
        double t = initvalue;
        for (int j = 0; j < cnt; ++j) {
            // executed at init on an empty network: out-link not full, in-link empty ⇒ nexttime = t + freeflowtime
            acts.push_back(HAct{n->turn_node[k], ACT_TURN, k, 0, -1, t + freeflow[n->turn_in[k]], t});
            t += 0.00000003;
        }
        initvalue += 0.00001;
    }
    for (int k = 0; k < n->n_odpairs; ++k) initvalue += 0.00001;
    for (int k = 0; k < n->n_dests; ++k) {
        for (int l = L - 1; l >= 0; --l) {  // dactions.insert(begin()) over ascending link ids ⇒ descending order
            if (n->link_out_node[l] != n->dest_node[k]) continue;
            for (int j = 0; j < n->link_lanes[l]; ++j)
                acts.push_back(HAct{n->dest_node[k], ACT_DEST, l, n->n_servers + k, n->dest_server[k],
                                    initvalue + freeflow[l], initvalue});  // empty link ⇒ next_action = t + freeflowtime
        }
        initvalue += 0.00001;
    }
    s->n_init_events = (long long)acts.size();
    // cross-node sharing of a stochastic server would serialise the whole network through one RNG stream
    {
        std::map<int, int> owner;
        for (const HAct &a : acts) {
            const int sv = a.kind == ACT_TURN ? n->turn_server[a.idx] : a.sv;
            if (sv < 0 || n->srv_type[sv] != 1) continue;
            auto it = owner.find(sv);
            if (it == owner.end()) owner[sv] = a.node;
            else
                MZ_REQUIRE(it->second == a.node, MZ_ELIMIT,
                           "stochastic server %d is shared by several nodes: its draws would have to follow the global "
                           "event order (SURVEY.md H5); give each node its own server or use deterministic servers", sv);
        }
    }
    // group by node (stable: keeps init order inside a node)
    std::vector<int> order(acts.size());
    for (size_t i = 0; i < acts.size(); ++i) order[i] = (int)i;
    std::stable_sort(order.begin(), order.end(), [&](int a, int b) { return acts[a].node < acts[b].node; });
    const int A = (int)acts.size();
    s->n_acts = A;
    d.n_acts = A;
    std::vector<int> act_off(N + 1, 0), kind(A), idx(A), aux(2 * (size_t)A);
    s->h_act_t.resize(A);
    s->h_act_tp.resize(A);
    s->h_act_node.resize(A);
    for (int i = 0; i < A; ++i) {
        const HAct &a = acts[order[i]];
        act_off[a.node + 1]++;
        kind[i] = a.kind;
        idx[i] = a.idx;
        aux[i] = a.aux;          // destination: slot of its server seed
        aux[A + i] = a.sv;       // destination: server index (or -1 for the node's own default server)
        s->h_act_t[i] = a.t;
        s->h_act_tp[i] = a.tp;
        s->h_act_node[i] = a.node;
    }
    for (int i = 0; i < N; ++i) act_off[i + 1] += act_off[i];

    // ---- node neighbours: nodes sharing a link
    std::vector<std::set<int>> nbs(N);
    for (int l = 0; l < L; ++l) {
        const int a = n->link_in_node[l], b = n->link_out_node[l];
        if (a != b) {
            nbs[a].insert(b);
            nbs[b].insert(a);
        }
    }
    std::vector<int> nb_off(N + 1, 0), nb;
    for (int i = 0; i < N; ++i) {
        nb.insert(nb.end(), nbs[i].begin(), nbs[i].end());
        nb_off[i + 1] = (int)nb.size();
    }

    // ---- trips per origin (global order preserved) and exit-log offsets
    std::vector<int> otrip_off(n->n_origins + 1, 0), otrips(n->n_trips);
    for (int v = 0; v < n->n_trips; ++v) otrip_off[n->trip_origin[v] + 1]++;
    for (int o = 0; o < n->n_origins; ++o) otrip_off[o + 1] += otrip_off[o];
    {
        std::vector<int> c(n->n_origins, 0);
        for (int v = 0; v < n->n_trips; ++v) otrips[otrip_off[n->trip_origin[v]] + c[n->trip_origin[v]]++] = v;
    }
    s->h_trip_log_off.assign(n->n_trips + 1, 0);
    for (int v = 0; v < n->n_trips; ++v)
        s->h_trip_log_off[v + 1] = s->h_trip_log_off[v] + (n->route_off[n->trip_route[v] + 1] - n->route_off[n->trip_route[v]]);
    s->total_log = s->h_trip_log_off[n->n_trips];

    // ---- initial node keys
    s->h_key_t.assign(N, INFINITY);
    s->h_key_tp.assign(N, INFINITY);
    for (int i = 0; i < N; ++i) {
        Key best{INFINITY, INFINITY, 0x7fffffff};
        for (int a = act_off[i]; a < act_off[i + 1]; ++a) {
            Key k{s->h_act_t[a], s->h_act_tp[a], a};
            if (key_less(k, best)) best = k;
        }
        if (node_type[i] == NODE_ORIGIN) {
            const int o = node_origin[i];
            if (otrip_off[o] < otrip_off[o + 1]) best = Key{n->trip_time[otrips[otrip_off[o]]], 0.0, i};
        }
        s->h_key_t[i] = best.t;
        s->h_key_tp[i] = best.tp;
    }

    // ---- upload
    int rc;
#define UP(field, v) if ((rc = dev_upload(s, d.field, v)) != MZ_OK) return rc
    UP(link_length, vec(n->link_length, L));
    UP(link_lanes, vec(n->link_lanes, L));
    UP(link_sdfunc, vec(n->link_sdfunc, L));
    UP(link_maxcap, maxcap);
    UP(link_qoff, qoff);
    UP(link_qcap, qcap);
    UP(link_freeflow, freeflow);
    UP(link_hist, s->h_link_hist);
    UP(sd_type, vec(n->sd_type, n->n_sdfuncs));
    UP(sd_vfree, vec(n->sd_vfree, n->n_sdfuncs));
    UP(sd_vmin, vec(n->sd_vmin, n->n_sdfuncs));
    UP(sd_romax, vec(n->sd_romax, n->n_sdfuncs));
    UP(sd_romin, vec(n->sd_romin, n->n_sdfuncs));
    UP(sd_alpha, vec(n->sd_alpha, n->n_sdfuncs));
    UP(sd_beta, vec(n->sd_beta, n->n_sdfuncs));
    UP(srv_type, vec(n->srv_type, n->n_servers));
    UP(srv_mu, vec(n->srv_mu, n->n_servers));
    UP(srv_sd, vec(n->srv_sd, n->n_servers));
    UP(srv_delay, vec(n->srv_delay, n->n_servers));
    UP(turn_server, vec(n->turn_server, n->n_turnings));
    UP(turn_in, vec(n->turn_in, n->n_turnings));
    UP(turn_out, vec(n->turn_out, n->n_turnings));
    UP(turn_lookback, vec(n->turn_lookback, n->n_turnings));
    UP(route_off, vec(n->route_off, (size_t)n->n_routes + 1));
    UP(route_links, vec(n->route_links, (size_t)n->route_off[n->n_routes]));
    UP(node_type, node_type);
    UP(node_act_off, act_off);
    UP(node_nb_off, nb_off);
    UP(node_nb, nb);
    UP(node_origin, node_origin);
    UP(origin_trip_off, otrip_off);
    UP(origin_trips, otrips);
    UP(trip_route, vec(n->trip_route, n->n_trips));
    UP(trip_time, vec(n->trip_time, n->n_trips));
    UP(trip_length, vec(n->trip_length, n->n_trips));
    UP(trip_log_off, s->h_trip_log_off);
    UP(act_kind, kind);
    UP(act_idx, idx);
    UP(act_aux, aux);
#undef UP
#define AL(field, cnt) if ((rc = dev_alloc(s, d.field, (size_t)(cnt))) != MZ_OK) return rc
    AL(q_head, L); AL(q_size, L); AL(nr_exits_blocked, L); AL(nr_passed, L); AL(tmp_passed, L); AL(curr_period, L);
    AL(blocked_until, L); AL(avg_time, L); AL(tmp_avg, L); AL(avgtimes, (size_t)L * P);
    AL(slab, slab);
    AL(veh_pos, n->n_trips); AL(veh_meters, n->n_trips); AL(veh_entry, n->n_trips); AL(veh_exit, n->n_trips);
    AL(arrival, n->n_trips); AL(exit_log, s->total_log);
    AL(act_t, A); AL(act_tp, A); AL(turn_blocked, n->n_turnings);
    s->n_servers_total = n->n_servers + n->n_dests;
    AL(srv_seed, s->n_servers_total);
    AL(origin_next, n->n_origins); AL(origin_vq_head, n->n_origins); AL(trip_parked, n->n_trips);
    AL(key_t[0], N); AL(key_t[1], N); AL(key_tp[0], N); AL(key_tp[1], N);
    AL(counters, 8);
#undef AL
    return sim_reset_state(s);
}

}  // namespace

extern "C" {

int mz_sim_create(int device, const MzNet *net, mz_sim **out)
{
    MZ_REQUIRE(out, MZ_EINVAL, "out handle pointer is null");
    *out = nullptr;
    if (int rc = sim_validate(net)) return rc;
    int ndev = 0;
    {
        cudaError_t e = cudaGetDeviceCount(&ndev);
        if (e != cudaSuccess || ndev == 0) {
            cudaGetLastError();
            mz::set_error("no CUDA device available (%s); mezzo_b200 has no CPU fallback",
                          e == cudaSuccess ? "device count 0" : cudaGetErrorString(e));
            return MZ_ENODEV;
        }
    }
    MZ_REQUIRE(device >= 0 && device < ndev, MZ_EINVAL, "device %d outside [0,%d)", device, ndev);
    MZ_CUDA(cudaSetDevice(device));
    mz_sim *s = new (std::nothrow) mz_sim;
    MZ_REQUIRE(s, MZ_ENOMEM, "out of host memory");
    s->device = device;
    cudaError_t e = cudaStreamCreateWithFlags(&s->s_own, cudaStreamNonBlocking);
    if (e == cudaSuccess) e = cudaEventCreate(&s->ev0);
    if (e == cudaSuccess) e = cudaEventCreate(&s->ev1);
    s->stream = s->s_own;
    int rc = e == cudaSuccess ? sim_build(s, net) : mz::cuda_fail(e, "stream/event creation", __FILE__, __LINE__);
    if (rc != MZ_OK) {
        mz_sim_destroy(s);
        return rc;
    }
    *out = s;
    return MZ_OK;
}

int mz_sim_reset(mz_sim *s)
{
    MZ_REQUIRE(s, MZ_EINVAL, "null handle");
    return sim_reset_state(s);
}

int mz_sim_run(mz_sim *s)
{
    MZ_REQUIRE(s, MZ_EINVAL, "null handle");
    MZ_REQUIRE(!s->ran, MZ_ESTATE, "mz_sim_run already executed; call mz_sim_reset first");
    MZ_CUDA(cudaSetDevice(s->device));
    auto w0 = std::chrono::steady_clock::now();
    const SimDev &d = s->dev;
    const int threads = 128;
    const int blocks = (d.n_nodes + threads - 1) / threads;
    unsigned long long *h_alive = nullptr;
    MZ_CUDA(cudaMallocHost((void **)&h_alive, sizeof(unsigned long long)));
    long long steps = 0, launches = 0;
    int parity = 0;
    MZ_CUDA(cudaEventRecord(s->ev0, s->stream));
    for (;;) {
        // a batch of windows; the "still work inside the horizon" flag is sampled on the last one
        for (int i = 0; i < s->check_every; ++i) {
            if (i == s->check_every - 1)
                MZ_CUDA(cudaMemsetAsync(d.counters + 4, 0, sizeof(unsigned long long), s->stream));
            k_sim_superstep<<<blocks, threads, 0, s->stream>>>(d, parity, -1);
            parity ^= 1;
            ++steps;
            ++launches;
        }
        MZ_CUDA(cudaMemcpyAsync(h_alive, d.counters + 4, sizeof(unsigned long long), cudaMemcpyDeviceToHost, s->stream));
        MZ_CUDA(cudaStreamSynchronize(s->stream));
        MZ_CUDA(cudaGetLastError());
        if (*h_alive == 0) break;
    }
    // "while (time < runtime) time = next()": the first event at/after the horizon still executes
    std::vector<double> kt(d.n_nodes), ktp(d.n_nodes);
    MZ_CUDA(cudaMemcpy(kt.data(), d.key_t[parity], sizeof(double) * d.n_nodes, cudaMemcpyDeviceToHost));
    MZ_CUDA(cudaMemcpy(ktp.data(), d.key_tp[parity], sizeof(double) * d.n_nodes, cudaMemcpyDeviceToHost));
    Key best{INFINITY, INFINITY, 0x7fffffff};
    for (int i = 0; i < d.n_nodes; ++i) {
        Key k{kt[i], ktp[i], i};
        if (key_less(k, best)) best = k;
    }
    double end_time = 0.0;
    if (best.id != 0x7fffffff && std::isfinite(best.t)) {
        k_sim_superstep<<<blocks, threads, 0, s->stream>>>(d, parity, best.id);
        parity ^= 1;
        ++launches;
        end_time = best.t;
    }
    MZ_CUDA(cudaEventRecord(s->ev1, s->stream));
    MZ_CUDA(cudaStreamSynchronize(s->stream));
    MZ_CUDA(cudaGetLastError());
    cudaFreeHost(h_alive);
    unsigned long long c[8];
    MZ_CUDA(cudaMemcpy(c, d.counters, sizeof(c), cudaMemcpyDeviceToHost));
    float ms = 0;
    MZ_CUDA(cudaEventElapsedTime(&ms, s->ev0, s->ev1));
    s->stats.n_traversals = (int64_t)c[0];
    s->stats.n_exits = (int64_t)c[1];
    s->stats.n_events = (int64_t)c[2] + s->n_init_events;
    s->stats.n_arrived = (int64_t)c[3];
    s->stats.n_supersteps = steps;
    s->stats.n_launches = launches;
    s->stats.kernel_ms = ms;
    s->stats.end_time = end_time;
    s->stats.total_ms = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - w0).count();
    s->ran = true;
    return MZ_OK;
}

int mz_sim_get_arrivals(mz_sim *s, double *arrival_time)
{
    MZ_REQUIRE(s && arrival_time, MZ_EINVAL, "null argument");
    MZ_CUDA(cudaSetDevice(s->device));
    if (s->n_trips) MZ_CUDA(cudaMemcpy(arrival_time, s->dev.arrival, sizeof(double) * s->n_trips, cudaMemcpyDefault));
    return MZ_OK;
}

int mz_sim_get_exit_log(mz_sim *s, MzExit *out, int64_t cap, int64_t *n_out)
{
    MZ_REQUIRE(s && n_out, MZ_EINVAL, "null argument");
    MZ_CUDA(cudaSetDevice(s->device));
    std::vector<double2> log((size_t)std::max<long long>(1, s->total_log));
    if (s->total_log)
        MZ_CUDA(cudaMemcpy(log.data(), s->dev.exit_log, sizeof(double2) * s->total_log, cudaMemcpyDeviceToHost));
    std::vector<int> route_off, route_links, trip_route;
    // link index per (trip, position) comes from the route table: fetch the static arrays back once
    const SimDev &d = s->dev;
    int n_routes_p1 = 0;
    {
        // route_off has n_routes+1 entries; its size is not stored, recover from the trips' routes
        trip_route.resize(std::max(1, s->n_trips));
        if (s->n_trips) MZ_CUDA(cudaMemcpy(trip_route.data(), d.trip_route, sizeof(int) * s->n_trips, cudaMemcpyDeviceToHost));
        int max_r = -1;
        for (int v = 0; v < s->n_trips; ++v) max_r = std::max(max_r, trip_route[v]);
        n_routes_p1 = max_r + 2;
        route_off.resize(std::max(1, n_routes_p1));
        if (max_r >= 0) {
            MZ_CUDA(cudaMemcpy(route_off.data(), d.route_off, sizeof(int) * n_routes_p1, cudaMemcpyDeviceToHost));
            route_links.resize(std::max(1, route_off[max_r + 1]));
            MZ_CUDA(cudaMemcpy(route_links.data(), d.route_links, sizeof(int) * route_off[max_r + 1], cudaMemcpyDeviceToHost));
        }
    }
    int64_t cnt = 0;
    for (int v = 0; v < s->n_trips; ++v) {
        const int r = trip_route[v];
        const long long b = s->h_trip_log_off[v], e = s->h_trip_log_off[v + 1];
        for (long long i = b; i < e; ++i) {
            if (std::isnan(log[i].y)) break;  // not exited (yet): later positions cannot have exits either
            if (out && cnt < cap) out[cnt] = MzExit{v + 1, route_links[route_off[r] + (int)(i - b)], log[i].x, log[i].y};
            ++cnt;
        }
    }
    *n_out = cnt;
    MZ_REQUIRE(!out || cnt <= cap, MZ_ELIMIT, "exit log holds %lld records, buffer %lld", (long long)cnt, (long long)cap);
    return MZ_OK;
}

int mz_sim_get_linktimes(mz_sim *s, double *avg)
{
    MZ_REQUIRE(s && avg, MZ_EINVAL, "null argument");
    MZ_CUDA(cudaSetDevice(s->device));
    MZ_CUDA(cudaMemcpy(avg, s->dev.avgtimes, sizeof(double) * (size_t)s->n_links * s->n_periods, cudaMemcpyDefault));
    return MZ_OK;
}

int mz_sim_last_stats(const mz_sim *s, mz_sim_stats *out)
{
    MZ_REQUIRE(s && out, MZ_EINVAL, "null argument");
    *out = s->stats;
    return MZ_OK;
}

int mz_sim_set_stream(mz_sim *s, void *cuda_stream)
{
    MZ_REQUIRE(s, MZ_EINVAL, "null handle");
    s->stream = cuda_stream ? (cudaStream_t)cuda_stream : s->s_own;
    return MZ_OK;
}

int mz_sim_destroy(mz_sim *s)
{
    if (!s) return MZ_OK;
    cudaSetDevice(s->device);
    for (void *p : s->allocs) cudaFree(p);
    if (s->ev0) cudaEventDestroy(s->ev0);
    if (s->ev1) cudaEventDestroy(s->ev1);
    if (s->s_own) cudaStreamDestroy(s->s_own);
    delete s;
    cudaGetLastError();
    return MZ_OK;
}
}
