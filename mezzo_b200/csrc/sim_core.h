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
// This header holds the data layout and every per-event function; it compiles both as CUDA device code (sim.cu, the
// product) and as plain C++ (tests/emu/sim_emu.cpp: a TEST-ONLY host emulation of the super-step schedule that lets the
// CPU test-suite check the window logic and the multi-rank hand-off without a GPU; the product never loads it).
#pragma once
#include <cmath>
#include <cstddef>
#include <cstdint>

#ifdef __CUDACC__
#define MZ_HD __device__
#define MZ_HDX __host__ __device__
#define MZ_ATOMIC_ADD(p, v) atomicAdd((p), (unsigned long long)(v))
#else
#define MZ_HD
#define MZ_HDX
#define MZ_ATOMIC_ADD(p, v) (*(p) += (unsigned long long)(v))
struct double2 { double x, y; };
static inline double2 make_double2(double x, double y) { double2 r; r.x = x; r.y = y; return r; }
#endif

#ifdef MZ_EMU_TRACE
void mz_emu_trace(double t, int kind, int idx);
#endif

namespace mzsim {

struct QEnt {
    double exit;
    int veh;
    int next;  // Vehicle::nextlink() cached at insertion (route successor of the link the entry sits in)
};
static_assert(sizeof(QEnt) == 16, "queue entry must be 16 bytes");

// Total order of events = the reference's multimap<double,Action*> with hint end(): by time, equal times FIFO by
// insertion. An action's event is inserted while its previous event executes, so FIFO among equal times = order of the
// PARENT executions: (t, t_parent, sub_parent), where sub is a Lamport counter of executions inside one time instant,
// maintained across neighbouring nodes (superstep_node). id is the node / action index, a last resort that only
// orders events which never interacted.
struct Key {
    double t, tp;
    int sub, id;
};
MZ_HDX inline bool key_less(const Key &a, const Key &b)
{
    if (a.t != b.t) return a.t < b.t;
    if (a.tp != b.tp) return a.tp < b.tp;
    if (a.sub != b.sub) return a.sub < b.sub;
    return a.id < b.id;
}

enum { ACT_TURN = 0, ACT_DEST = 1, ACT_OD = 2 };
enum { NODE_JUNCTION = 0, NODE_ORIGIN = 1, NODE_DEST = 2 };

// everything the kernels need, passed by value
struct SimDev {
    int n_nodes, n_links, n_turnings, n_trips, n_periods, n_acts;
    int max_events;  // cap on events one node executes per window (bounds the window's latency; never changes results)
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
    const int *node_act_off, *node_nb_off, *node_nb;
    const int *origin_trip_off, *origin_trips;  // trips per origin in ODaction execution order (virtual-queue order)
    const int *od_off, *od_trips;               // trips per OD pair (one ODaction each), in order
    const int *trip_slot;                       // position of a trip inside its origin's list
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
    int *act_sub;          // Lamport sub-counter of the action's last execution (orders equal-time bookings)
    double *node_last_t;   // time instant of the node's most recent execution
    int *node_last_sub;    // largest sub-counter the node used at that instant
    const int *act_kind, *act_idx, *act_aux;
    unsigned char *turn_blocked;
    long long *srv_seed;
    // origins
    int *od_next, *origin_vq_head;
    unsigned char *trip_parked;
    // published node keys (double buffered)
    double *key_t[2], *key_tp[2];
    int *key_sub[2];
    // counters: [0] traversals [1] exits [2] events [3] arrived [4] alive flag [5] error flag
    unsigned long long *counters;
};

// ------------------------------------------------------------------------------------------------ Random / Server
MZ_HD inline double urandom(long long &seed)  // B/Random.cpp:85-98
{
    const long long M = 2147483647LL, A = 48271LL, Q = M / A, R = M % A;
    long long s = A * (seed % Q) - R * (seed / Q);
    s = (s > 0) ? s : (s + M);
    seed = s;
    return (double)s / (double)M;
}
MZ_HD inline double nrandom(long long &seed, double mean, double sd)  // B/Random.cpp:126-132
{
    const double r1 = urandom(seed), r2 = urandom(seed);
    const double r = -2.0 * log(r1);
    if (r > 0.0) return mean + sd * sqrt(r) * sin((3.14159265358979323846 * 2.0) * r2);
    return mean;
}
MZ_HD inline double server_next(int type, double mu, double sd, double delay, long long &seed, double t)
{
    if (type == 0) return t;               // ConstServer   B/server.h:74
    if (type == 2) return t + mu + delay;  // DetServer     B/server.h:98
    const double x = nrandom(seed, mu, sd);  // Server::next  B/server.cpp:46-48
    return t + (x > 0.1 ? x : 0.1);
}
MZ_HD inline double sd_speed(const SimDev &d, int f, double ro)  // B/sdfunc.cpp:12-19, 45-53
{
    if (ro <= d.sd_romin[f]) return d.sd_vfree[f];
    if (ro > d.sd_romax[f]) return d.sd_vmin[f];
    const double x = (ro - d.sd_romin[f]) / (d.sd_romax[f] - d.sd_romin[f]);
    if (d.sd_type[f] == 0) return d.sd_vmin[f] + (d.sd_vfree[f] - d.sd_vmin[f]) * (1 - x);
    return d.sd_vmin[f] + (d.sd_vfree[f] - d.sd_vmin[f]) * pow((1 - pow(x, d.sd_alpha[f])), d.sd_beta[f]);
}

// ------------------------------------------------------------------------------------------------ warp abstraction
// One WARP cooperates on one node: every lane runs the same control flow on the same (uniform) scalars, queue scans
// and shifts are spread over the lanes (coalesced 16-byte entries, ballot to locate), and lane 0 alone stores.
// The host emulation compiles the same code with a "warp" of one lane.
#ifdef __CUDACC__
#define MZ_W 32
MZ_HD inline int mz_lane() { return (int)(threadIdx.x & 31u); }
MZ_HD inline unsigned mz_ballot(bool p) { return __ballot_sync(0xffffffffu, p); }
MZ_HD inline void mz_syncwarp() { __syncwarp(); }
MZ_HD inline int mz_ffs(unsigned b) { return __ffs((int)b); }
MZ_HD inline int mz_clz(unsigned b) { return __clz((int)b); }
MZ_HD inline double mz_shfl_xor(double v, int m) { return __shfl_xor_sync(0xffffffffu, v, m); }
MZ_HD inline int mz_shfl_xor(int v, int m) { return __shfl_xor_sync(0xffffffffu, v, m); }
#else
#define MZ_W 1
inline int mz_lane() { return 0; }
inline unsigned mz_ballot(bool p) { return p ? 1u : 0u; }
inline void mz_syncwarp() {}
inline int mz_ffs(unsigned b) { return __builtin_ffs((int)b); }
inline int mz_clz(unsigned b) { return __builtin_clz(b); }
inline double mz_shfl_xor(double v, int) { return v; }
inline int mz_shfl_xor(int v, int) { return v; }
#endif
template <class T>
MZ_HD inline void wr(T *p, T v)  // uniform store: lane 0 only
{
    if (mz_lane() == 0) *p = v;
}
// warp-wide minimum of a key (every lane ends up with the same result)
MZ_HD inline Key key_warp_min(Key k)
{
    for (int m = MZ_W / 2; m > 0; m >>= 1) {
        Key o;
        o.t = mz_shfl_xor(k.t, m);
        o.tp = mz_shfl_xor(k.tp, m);
        o.sub = mz_shfl_xor(k.sub, m);
        o.id = mz_shfl_xor(k.id, m);
        if (key_less(o, k)) k = o;
    }
    return k;
}

// ------------------------------------------------------------------------------------------------ queue (ring buffer in list order)
struct QRef {
    QEnt *base;
    int cap, head, size;
};
MZ_HD inline QRef q_open(const SimDev &d, int l)
{
    QRef q;
    q.base = d.slab + d.link_qoff[l];
    q.cap = d.link_qcap[l];
    q.head = d.q_head[l];
    q.size = d.q_size[l];
    return q;
}
MZ_HD inline void q_close(const SimDev &d, int l, const QRef &q)
{
    wr(d.q_head + l, q.head);
    wr(d.q_size + l, q.size);
}
MZ_HD inline QEnt &q_at(const QRef &q, int i)
{
    int p = q.head + i;
    if (p >= q.cap) p -= q.cap;
    return q.base[p];
}
// smallest list position whose entry heads for link `next`, or size
MZ_HD inline int q_find_first_next(const QRef &q, int next)
{
    for (int base = 0; base < q.size; base += MZ_W) {
        const int i = base + mz_lane();
        const unsigned b = mz_ballot(i < q.size && q_at(q, i).next == next);
        if (b) return base + mz_ffs(b) - 1;
    }
    return q.size;
}
// smallest list position whose exit_time > t, or size  (Q::nr_running = size - this, B/q.h:67-68)
MZ_HD inline int q_find_first_running(const QRef &q, double t)
{
    for (int base = 0; base < q.size; base += MZ_W) {
        const int i = base + mz_lane();
        const unsigned b = mz_ballot(i < q.size && q_at(q, i).exit > t);
        if (b) return base + mz_ffs(b) - 1;
    }
    return q.size;
}
// largest list position whose exit_time is (gt ? > t : <= t), or -1
MZ_HD inline int q_find_last(const QRef &q, double t, bool gt)
{
    for (int base = ((q.size - 1) / MZ_W) * MZ_W; base >= 0 && q.size > 0; base -= MZ_W) {
        const int i = base + mz_lane();
        bool p = false;
        if (i < q.size) {
            const double e = q_at(q, i).exit;
            p = gt ? (e > t) : (e <= t);
        }
        const unsigned b = mz_ballot(p);
        if (b) return base + (31 - mz_clz(b));
    }
    return -1;
}
// entries [from, to) move one slot towards the tail (to [from+1, to+1)); chunks run from the tail so nothing unread is hit
MZ_HD inline void q_shift_up(const QRef &q, int from, int to)
{
    mz_syncwarp();
    for (int hi = to; hi > from; hi -= MZ_W) {
        const int i = hi - 1 - mz_lane();
        QEnt tmp;
        const bool act = i >= from;
        if (act) tmp = q_at(q, i);
        mz_syncwarp();
        if (act) q_at(q, i + 1) = tmp;
        mz_syncwarp();
    }
}
// entries [from+1, to+1) move one slot towards the head (to [from, to))
MZ_HD inline void q_shift_down(const QRef &q, int from, int to)
{
    mz_syncwarp();
    for (int lo = from; lo < to; lo += MZ_W) {
        const int i = lo + mz_lane();
        QEnt tmp;
        const bool act = i < to;
        if (act) tmp = q_at(q, i + 1);
        mz_syncwarp();
        if (act) q_at(q, i) = tmp;
        mz_syncwarp();
    }
}
// erase list position k: the k entries in front of it move one slot towards the tail, the head advances
MZ_HD inline void q_erase(QRef &q, int k)
{
    q_shift_up(q, 0, k);
    q.head = q.head + 1 == q.cap ? 0 : q.head + 1;
    q.size--;
}
// insert at list position p (0..size)
MZ_HD inline void q_insert(QRef &q, int p, const QEnt &e)
{
    if (p >= q.size - p) {  // closer to the tail: shift the tail part back by one
        q_shift_up(q, p, q.size);
        q.size++;
    } else {  // closer to the head: move the head one slot forward and shift the front part
        q.head = q.head == 0 ? q.cap - 1 : q.head - 1;
        q.size++;
        q_shift_down(q, 0, p);
    }
    if (mz_lane() == 0) q_at(q, p) = e;
    mz_syncwarp();
}

MZ_HD inline int veh_next_after(const SimDev &d, int v, int pos)
{
    const int r = d.trip_route[v];
    const int i = d.route_off[r] + pos + 1;
    return i < d.route_off[r + 1] ? d.route_links[i] : -1;
}

// ------------------------------------------------------------------------------------------------ Link
MZ_HD inline bool link_full_t(const SimDev &d, int l, double t)  // Link::full(time) B/link.cpp:155-183
{
    const double bu = d.blocked_until[l];
    if (bu == -2.0) return true;
    if (d.q_size[l] >= d.link_maxcap[l]) {
        if (d.nr_exits_blocked[l] > 0) wr(d.blocked_until + l, -2.0);
        return true;
    }
    if (bu == -1.0) return false;
    if (bu <= t) {
        wr(d.blocked_until + l, -1.0);
        return false;
    }
    return true;
}

// Link::enter_veh (B/link.cpp:418-523) = density_running_only (670-687) + Sdfunc::speed + Q::enter_veh (B/q.cpp:52-72)
MZ_HD inline void link_enter_veh(const SimDev &d, int l, int v, int pos_before, double t, unsigned long long &n_trav)
{
    mz_syncwarp();
    QRef q = q_open(d, l);
    // nr_running: elements from the first one with exit_time > t to the end (B/q.h:67-68)
    const int nr = q.size - q_find_first_running(q, t);
    // calc_space: vehicle lengths from the back while exit_time <= t (B/q.cpp:346-361); 0 unless the tail is queued
    double space = 0.0;
    if (q.size > 0) {
        const int last_running = q_find_last(q, t, true);
        for (int i = q.size - 1; i > last_running; --i) space += d.trip_length[q_at(q, i).veh];
    }
    const int lanes = d.link_lanes[l], len = d.link_length[l];
    const double queue_space = space / lanes;
    const double running_length = len - queue_space;
    double ro = 0.0;
    if (nr && running_length) ro = nr / (running_length * (lanes / 1000.0));
    const double speed = sd_speed(d, d.link_sdfunc[l], ro);
    const double exit_time = t + (len / speed);
    const int pos = pos_before + 1;
    wr(d.veh_pos + v, pos);
    wr(d.veh_entry + v, t);
    QEnt e;
    e.exit = exit_time;
    e.veh = v;
    e.next = veh_next_after(d, v, pos);
    if (q.size == 0) {
        q_insert(q, 0, e);
    } else {
        // "while (viter->first > ttime && viter != begin) --viter; insert(++viter)": behind the last entry that is not
        // later than the newcomer, but never in front of the head
        int i = q_find_last(q, exit_time, false);
        if (i < 0) i = 0;
        q_insert(q, i + 1, e);
    }
    q_close(d, l, q);
    mz_syncwarp();
    ++n_trav;
}

// bookkeeping of a successful Link::exit_veh (B/link.cpp:536-581 / 601-648) + exit log
MZ_HD inline void link_exit_bookkeeping(const SimDev &d, int l, int v, int pos, double t, unsigned long long &n_exits)
{
    const double entrytime = d.veh_entry[v], traveltime = t - entrytime;
    const int np = d.nr_passed[l];
    wr(d.avg_time + l, (np * d.avg_time[l] + traveltime) / (np + 1));
    const int cp = d.curr_period[l];
    if ((cp + 1) * d.periodlength < entrytime) {
        double ta = d.tmp_avg[l];
        if (ta == 0.0) ta = cp < d.n_periods ? d.link_hist[(size_t)l * d.n_periods + cp] : 0.0;
        if (cp < d.n_periods) wr(d.avgtimes + (size_t)l * d.n_periods + cp, ta);
        wr(d.curr_period + l, cp + 1);
        wr(d.tmp_avg + l, 0.0);
        wr(d.tmp_passed + l, 0);
    } else {
        const int tp = d.tmp_passed[l];
        wr(d.tmp_avg + l, (tp * d.tmp_avg[l] + traveltime) / (tp + 1));
        wr(d.tmp_passed + l, tp + 1);
    }
    wr(d.nr_passed + l, np + 1);
    wr(d.veh_meters + v, d.veh_meters[v] + d.link_length[l]);
    wr(d.exit_log + (d.trip_log_off[v] + pos), make_double2(entrytime, t));
    ++n_exits;
}

// Link::update_exit_times + Q::update_exit_times (B/link.cpp:320-390, B/q.cpp:74-129); rare (only on unblocking)
MZ_HD inline void link_update_exit_times(const SimDev &d, int l, double t, int next, int lookback)
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
    const QRef q = q_open(d, l);
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
                wr(&e.exit, newtime);
                t0 = newtime;
            }
            i++;
        } else {
            cnt++;
            const double newtime = t0 + vl / v_sw + vl / v_exit;
            const int lanes = d.link_lanes[l];
            for (int k = 0; k < lanes && i < q.size; ++k) {
                wr(&q_at(q, i).exit, newtime);
                i++;
            }
            t0 = newtime;
        }
    }
    if (d.blocked_until[l] == -2.0) wr(d.blocked_until + l, (d.link_length[l] / v_sw) + t);
    mz_syncwarp();
}

struct Tally {
    unsigned long long trav = 0, exits = 0, events = 0, arrived = 0;
};

// TurnAction::execute + Turning::process_veh (B/turning.cpp:237-265, 45-146); returns the re-booking time
MZ_HD inline double turn_execute(const SimDev &d, int k, double t, Tally &ty)
{
    const int in = d.turn_in[k], out = d.turn_out[k], sv = d.turn_server[k];
    double nexttime = t;
    bool ok = false;
    if (link_full_t(d, out, t)) {
        if (!d.turn_blocked[k]) {
            wr(d.turn_blocked + k, (unsigned char)1);
            wr(d.nr_exits_blocked + in, d.nr_exits_blocked[in] + 1);
        }
        nexttime = t + 1.0;
    } else {
        if (d.turn_blocked[k]) {
            wr(d.turn_blocked + k, (unsigned char)0);
            link_update_exit_times(d, in, t, out, d.turn_lookback[k]);
            wr(d.nr_exits_blocked + in, d.nr_exits_blocked[in] - 1);
        }
        QRef q = q_open(d, in);
        if (q.size == 0) {
            nexttime = t + d.link_freeflow[in];
        } else {
            // Q::exit_veh(time, nextlink, lookback): first vehicle in LIST order heading for `out` (B/q.cpp:156-222)
            const int n = q_find_first_next(q, out);
            if (n == q.size) {
                nexttime = t + d.link_freeflow[in];  // next_action = time + freeflowtime
            } else {
                const QEnt e = q_at(q, n);
                if (e.exit <= t) {
                    if (n < d.turn_lookback[k]) {
                        q_erase(q, n);
                        q_close(d, in, q);
                        const int pos = d.veh_pos[e.veh];
                        link_exit_bookkeeping(d, in, e.veh, pos, t, ty.exits);
                        link_enter_veh(d, out, e.veh, pos, t + d.srv_delay[sv], ty.trav);
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
        wr(d.srv_seed + sv, seed);
        return nt;
    }
    return nexttime < t ? t + 0.1 : nexttime;
}

// Daction::execute (B/node.cpp:637-673) with Q::exit_veh(time) (B/q.cpp:227-249)
MZ_HD inline double dest_execute(const SimDev &d, int dst_server_slot, int sv, int l, double t, Tally &ty)
{
    QRef q = q_open(d, l);
    if (q.size == 0) return t + d.link_freeflow[l];
    const QEnt e = q_at(q, 0);
    if (e.exit <= t) {
        q_erase(q, 0);
        q_close(d, l, q);
        link_exit_bookkeeping(d, l, e.veh, d.veh_pos[e.veh], t, ty.exits);
        wr(d.arrival + e.veh, t);  // Vehicle::report(time)
        ++ty.arrived;
        long long seed = d.srv_seed[dst_server_slot];
        double nt;
        if (sv >= 0) nt = server_next(d.srv_type[sv], d.srv_mu[sv], d.srv_sd[sv], d.srv_delay[sv], seed, t);
        else nt = server_next(1, 0.1, 0.5, 0.001, seed, t);  // Destination's own Server(-20000,0,0.1,0.5,0.001)
        wr(d.srv_seed + dst_server_slot, seed);
        return nt;
    }
    return e.exit;
}

// Origin::insert_veh / transfer_veh with the InputLink virtual queue (B/node.cpp:126-250, B/link.cpp:949-995).
// The virtual queue of origin o is the set of its trips in [vq_head, slot) whose `parked` flag is set, in trip order.
MZ_HD inline void origin_insert(const SimDev &d, int o, int slot, double t, Tally &ty)
{
    const int *trips = d.origin_trips;
    const int v = trips[slot];
    const int first = d.route_links[d.route_off[d.trip_route[v]]];
    int vqh = d.origin_vq_head[o];
    while (vqh < slot && !d.trip_parked[trips[vqh]]) ++vqh;  // drop leading non-parked trips
    const bool vq_empty = vqh == slot;
    if (d.q_size[first] >= d.link_maxcap[first]) {  // link->full(): park
        wr(d.trip_parked + v, (unsigned char)1);
        wr(d.veh_entry + v, t);
    } else if (vq_empty) {
        link_enter_veh(d, first, v, -1, t, ty.trav);
    } else {
        wr(d.trip_parked + v, (unsigned char)1);
        wr(d.veh_entry + v, t);
        mz_syncwarp();
        // transfer_veh: head-first, every parked vehicle whose first link is `first`, while the link is not full
        for (int i = vqh; i <= slot; ++i) {
            if (d.q_size[first] >= d.link_maxcap[first]) break;
            const int w = trips[i];
            if (!d.trip_parked[w]) continue;
            if (d.route_links[d.route_off[d.trip_route[w]]] != first) continue;
            wr(d.trip_parked + w, (unsigned char)0);
            link_enter_veh(d, first, w, -1, t, ty.trav);
        }
    }
    wr(d.origin_vq_head + o, vqh);
}

// ------------------------------------------------------------------------------------------------ the super-step
// earliest pending action of node n (lanes scan the node's action list); `which` = -1 if the node has none
MZ_HD inline Key node_min_key(const SimDev &d, int n, int &which)
{
    Key best;
    best.t = INFINITY;
    best.tp = INFINITY;
    best.sub = 0x7fffffff;
    best.id = 0x7fffffff;
    for (int a = d.node_act_off[n] + mz_lane(); a < d.node_act_off[n + 1]; a += MZ_W) {
        Key k;
        k.t = d.act_t[a];
        k.tp = d.act_tp[a];
        k.sub = d.act_sub[a];
        k.id = a;
        if (key_less(k, best)) best = k;
    }
    best = key_warp_min(best);
    which = best.id == 0x7fffffff ? -1 : best.id;
    return best;
}

// executes the earliest event of node n (identified by `which`); re-books that action
MZ_HD inline void node_execute(const SimDev &d, int which, double t, int sub, Tally &ty)
{
    ++ty.events;
#ifdef MZ_EMU_TRACE
    mz_emu_trace(t, d.act_kind[which], d.act_idx[which]);
#endif
    double nt;
    const int kind = d.act_kind[which];
    if (kind == ACT_TURN) {
        nt = turn_execute(d, d.act_idx[which], t, ty);
    } else if (kind == ACT_DEST) {
        nt = dest_execute(d, d.act_aux[which], d.act_aux[which + d.n_acts], d.act_idx[which], t, ty);
    } else {  // ODaction::execute (B/od.cpp:69-108): the pair's next pre-generated trip enters at its origin
        const int k = d.act_idx[which];
        const int pos = d.od_off[k] + d.od_next[k];
        const int v = d.od_trips[pos];
        origin_insert(d, d.act_aux[which], d.origin_trip_off[d.act_aux[which]] + d.trip_slot[v], t, ty);
        wr(d.od_next + k, d.od_next[k] + 1);
        nt = pos + 1 < d.od_off[k + 1] ? d.trip_time[d.od_trips[pos + 1]] : INFINITY;
    }
    wr(d.act_tp + which, t);
    wr(d.act_sub + which, sub);
    wr(d.act_t + which, nt);
    mz_syncwarp();
}

// One conservative window of node n, executed by one warp (host emulation: one lane).
// `final_node` >= 0 selects the clean-up launch that executes exactly one event (the first one at/after the horizon).
MZ_HD inline void superstep_node(const SimDev &d, int n, int parity, int final_node)
{
    const double *kt = d.key_t[parity], *ktp = d.key_tp[parity];
    const int *ks = d.key_sub[parity];
    Key mine;
    mine.t = kt[n];
    mine.tp = ktp[n];
    mine.sub = ks[n];
    mine.id = n;
    Tally ty;
    bool touched = false;
    if (final_node >= 0 ? n == final_node : mine.t < d.runtime) {
        // the window: everything before the earliest pending key of any neighbour. While scanning the neighbours also
        // pick up the latest time instant at which one of them executed and the largest Lamport sub-counter used there
        // (no neighbour executes during this super-step, so these are stable).
        Key bound;
        bound.t = INFINITY;
        bound.tp = INFINITY;
        bound.sub = 0x7fffffff;
        bound.id = 0x7fffffff;
        Key nbl;  // (last instant, -sub) folded into a key so one warp reduction yields the latest instant and its max sub
        nbl.t = INFINITY;
        nbl.tp = 0.0;
        nbl.sub = 0;
        nbl.id = 0;
        for (int i = d.node_nb_off[n] + mz_lane(); i < d.node_nb_off[n + 1]; i += MZ_W) {
            const int m = d.node_nb[i];
            Key k;
            k.t = kt[m];
            k.tp = ktp[m];
            k.sub = ks[m];
            k.id = m;
            if (key_less(k, bound)) bound = k;
            Key l;
            l.t = -d.node_last_t[m];
            l.tp = 0.0;
            l.sub = -d.node_last_sub[m];
            l.id = 0;
            if (key_less(l, nbl)) nbl = l;
        }
        bound = key_warp_min(bound);
        nbl = key_warp_min(nbl);
        const double nb_t = -nbl.t;
        const int nb_sub = -nbl.sub;
        if (final_node >= 0 || key_less(mine, bound)) {
            double last_t = d.node_last_t[n];
            int last_sub = d.node_last_sub[n];
            for (int it = 0;; ++it) {
                int which;
                Key k = node_min_key(d, n, which);
                if (which < 0) break;
                if (final_node >= 0) {
                    if (it == 1) break;  // exactly one event: the first one at/after the horizon
                } else {
                    if (it >= d.max_events) break;
                    if (!(k.t < d.runtime)) break;
                    Key kn = k;
                    kn.id = n;  // keys are compared across nodes by node index
                    if (!key_less(kn, bound)) break;
                }
                // Lamport counter inside the time instant k.t: one more than anything this node or a neighbour
                // has executed at exactly this time
                int sub = 0;
                if (last_t == k.t) sub = last_sub + 1;
                if (nb_t == k.t && nb_sub + 1 > sub) sub = nb_sub + 1;
                last_t = k.t;
                last_sub = sub;
                node_execute(d, which, k.t, sub, ty);
                touched = true;
            }
            if (touched) {
                wr(d.node_last_t + n, last_t);
                wr(d.node_last_sub + n, last_sub);
            }
        }
    }
    Key k = mine;
    if (touched) {
        int which;
        k = node_min_key(d, n, which);
    }
    if (mz_lane() == 0) {
        d.key_t[parity ^ 1][n] = k.t;
        d.key_tp[parity ^ 1][n] = k.tp;
        d.key_sub[parity ^ 1][n] = k.sub;
        if (k.t < d.runtime) d.counters[4] = 1;  // somebody still has work inside the horizon
        if (ty.events) {
            MZ_ATOMIC_ADD(d.counters + 2, ty.events);
            if (ty.trav) MZ_ATOMIC_ADD(d.counters + 0, ty.trav);
            if (ty.exits) MZ_ATOMIC_ADD(d.counters + 1, ty.exits);
            if (ty.arrived) MZ_ATOMIC_ADD(d.counters + 3, ty.arrived);
        }
    }
}

}  // namespace mzsim
