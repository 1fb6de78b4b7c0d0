/*
 * TEST INFRASTRUCTURE — CPU restatement ("port") of the reference's link-update event loop (hot path 1).
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may load this.
 *
 * What it restates (B/ = /root/reference/branches/busmezzo/mezzo_lib/src/):
 *   Network::init / Network::step          B/network.cpp:7657-7754, 7792-7821     (first events, run loop, H8)
 *   Eventlist (FIFO among equal times)     B/eventlist.h:58-71
 *   TurnAction::execute, Turning::process_veh  B/turning.cpp:237-265, 45-146
 *   Link::full / enter_veh / exit_veh / update_exit_times / density_running_only / next_action
 *                                          B/link.cpp:155-183, 418-523, 528-655, 320-390, 670-687, 188-197
 *   Q::enter_veh / exit_veh (both) / update_exit_times / calc_space / nr_running
 *                                          B/q.cpp:52-72, 156-249, 74-129, 346-361, B/q.h:67-68
 *   Sdfunc::speed, DynamitSdfunc::speed    B/sdfunc.cpp:12-19, 45-53
 *   Server::next, ConstServer, DetServer   B/server.cpp:32-50, B/server.h:65-99
 *   Random::urandom / nrandom(mean,sd)     B/Random.cpp:85-98, 126-132
 *   Daction::execute / process_veh         B/node.cpp:637-673
 *   Origin::insert_veh / transfer_veh, InputLink   B/node.cpp:126-250, B/link.cpp:949-995
 *   Bus at a stop link / leaving a stop    B/link.cpp:489-516, B/busline.cpp:1196-1263 (transit hooks, row f3: the dwell
 *                                          time of a visit is an INPUT here — the transit model itself is not on the path;
 *                                          pinned on the reference's BusMezzo fixture, tests/golden/bus_sf_golden.json)
 * Trip generation (ODaction) is a host pre-pass: trips arrive here as a time-ordered table.
 *
 * Parity pinning: the reference's tests pin nothing for car traffic, so this restatement is pinned
 * against the reference ITSELF: oracle/_ref/mezzo_ref_run (unmodified mezzo_lib compiled in place, exit log taken by
 * a compile-time macro) on generated networks — tests/test_oracle_sim.py (live, when /root/reference built _ref) and
 * the committed fixtures tests/golden/sim_*.npz produced by tests/golden/make_sim_golden.py.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#include "../include/mezzo_b200.h"

typedef struct { double exit; int32_t veh; } QEnt; /* Veh_in_Q, B/q.h:38 */

typedef struct {
    QEnt *q; int32_t size, cap;          /* std::list in list order */
    int32_t maxcap; double freeflow;     /* B/link.cpp:11, 42-45 */
    double blocked_until; int32_t nr_exits_blocked;
    double q_next_action;                /* Q::next_action */
    int q_ok;
    double avg_time, tmp_avg; int32_t nr_passed, tmp_passed, curr_period;
    int32_t sd_cur, sd_tmp;              /* Link::sdfunc / temp_sdfunc (incidents, B/link.cpp:854-868) */
} OLink;

typedef struct { int32_t route, pos; double entry, exit, length, start; int32_t meters; } OVeh;

typedef struct { double t, tp, tpp; int64_t seq; int32_t act; } Ev;

enum { ACT_TURN = 0, ACT_DEST = 1, ACT_TRIPS = 2, ACT_SIGC = 3, ACT_SIGP = 4, ACT_SIGS = 5, ACT_INC = 6, ACT_BUSDEP = 7 };
typedef struct { int32_t kind, idx, aux; } Act;

typedef struct {
    const MzNet *n;
    OLink *lk; OVeh *veh;
    int64_t *srv_seed;       /* one Random per server, + one per destination default server */
    uint8_t *turn_blocked;
    /* traffic signals (B/trafficsignal.cpp): Turning::active, the turnings' action ids, SignalControl / SignalPlan / Stage state */
    uint8_t *turn_active, *sigc_active, *sigp_active, *sigs_active;
    int32_t *turn_act0, *turn_nact, *sigc_cur, *sigp_cur, *sigc_act, *sigp_act, *sigs_act;
    /* virtual queues (InputLink) per origin */
    int32_t **vq; int32_t *vq_size, *vq_cap;
    Act *acts; int32_t n_acts;
    Ev *heap; int64_t heap_n, heap_cap, seq;
    int32_t *od_off, *od_trips, *od_next; /* trips per OD pair (ODaction), in order */
    double *avgtimes;        /* [L*P] */
    double *arrival;         /* [n_trips] */
    MzExit *log; int64_t log_n, log_cap;
    /* transit overlay hooks (row f3): stops already booked per bus, the bus's departure action, the visit log, and — the
       stand-in for the host's transit model, which is NOT part of the path — a dwell time per stop visit */
    int32_t *bus_next, *bus_act; const double *dwell;
    mz_stop_visit *visits; int64_t n_visits;
    mz_sim_stats st;
} OSim;

/* ---------------- Random (Park–Miller variant, B/Random.cpp:85-98) ---------------- */
static double urandom(int64_t *seed)
{
    const int64_t M = 2147483647, A = 48271, Q = M / A, R = M % A;
    int64_t s = *seed;
    s = A * (s % Q) - R * (s / Q);
    s = (s > 0) ? s : (s + M);
    *seed = s;
    return (double)s / (double)M;
}
static double nrandom(int64_t *seed, double mean, double sd) /* B/Random.cpp:126-132 */
{
    double r1 = urandom(seed), r2 = urandom(seed);
    double r = -2.0 * log(r1);
    if (r > 0.0) return mean + sd * sqrt(r) * sin((3.14159265358979323846 * 2.0) * r2) /* TWO_PI = M_PI*2.0, B/MMath.h:35 */;
    return mean;
}

/* ---------------- heap keyed (time, insertion seq) = multimap with hint end() ---------------- */
/* Exact reference order is (time, insertion sequence). orc_tiebreak != 0 switches to orderings that need no global
 * sequence counter, to evaluate the device's tie rule on the CPU: 1 = (t, t_prev, act), 2 = (t, t_prev, t_prevprev, act) */
int orc_tiebreak = 0;
static int ev_less(const Ev *a, const Ev *b)
{
    if (a->t != b->t) return a->t < b->t;
    if (orc_tiebreak == 0) return a->seq < b->seq;
    if (a->tp != b->tp) return a->tp < b->tp;
    if (orc_tiebreak == 2 && a->tpp != b->tpp) return a->tpp < b->tpp;
    return a->act < b->act;
}
static double g_tp = 0.0, g_tpp = 0.0; /* booking context of the next heap_push */
static void heap_push(OSim *s, double t, int32_t act)
{
    if (s->heap_n == s->heap_cap) {
        s->heap_cap = s->heap_cap ? 2 * s->heap_cap : 1024;
        s->heap = (Ev *)realloc(s->heap, sizeof(Ev) * (size_t)s->heap_cap);
    }
    int64_t i = s->heap_n++;
    Ev e = { t, g_tp, g_tpp, s->seq++, act };
    while (i > 0) {
        int64_t p = (i - 1) / 2;
        if (!ev_less(&e, &s->heap[p])) break;
        s->heap[i] = s->heap[p];
        i = p;
    }
    s->heap[i] = e;
}
static Ev heap_pop(OSim *s)
{
    Ev top = s->heap[0], last = s->heap[--s->heap_n];
    int64_t i = 0;
    for (;;) {
        int64_t c = 2 * i + 1;
        if (c >= s->heap_n) break;
        if (c + 1 < s->heap_n && ev_less(&s->heap[c + 1], &s->heap[c])) ++c;
        if (!ev_less(&s->heap[c], &last)) break;
        s->heap[i] = s->heap[c];
        i = c;
    }
    if (s->heap_n > 0) s->heap[i] = last;
    return top;
}

/* ---------------- Sdfunc ---------------- */
static double sd_speed(const MzNet *n, int32_t f, double ro)
{
    if (ro <= n->sd_romin[f]) return n->sd_vfree[f];
    if (ro > n->sd_romax[f]) return n->sd_vmin[f];
    double x = (ro - n->sd_romin[f]) / (n->sd_romax[f] - n->sd_romin[f]);
    if (n->sd_type[f] == 0) return n->sd_vmin[f] + (n->sd_vfree[f] - n->sd_vmin[f]) * (1 - x);
    return n->sd_vmin[f] + (n->sd_vfree[f] - n->sd_vmin[f]) * pow((1 - pow(x, n->sd_alpha[f])), n->sd_beta[f]);
}

/* ---------------- Server::next ---------------- */
static double lnrandom(int64_t *seed, double m, double v) /* Random::lnrandom, B/Random.cpp:151-157 */
{
    double mu_ = log(m * m / sqrt((m * m) + (v * v)));
    double sd_ = sqrt(log(1 + (v * v) / (m * m)));
    return exp(nrandom(seed, mu_, sd_));
}

static double loglogisticrandom(int64_t *seed, double alpha, double beta) /* Random::loglogisticrandom, B/Random.cpp:165-172 */
{
    double alpha_logistic = log(alpha);
    double beta_logistic = 1 / beta;
    /* log(urandom()/(1-urandom())): two draws in one expression; the compiled reference (g++ -O2) takes the numerator's
     * draw first — pinned by the loglogistic golden cases */
    double u_num = urandom(seed);
    double u_den = urandom(seed);
    double logistic = alpha_logistic + beta_logistic * log(u_num / (1 - u_den));
    return exp(logistic);
}

static double server_next(OSim *s, int32_t type, double mu, double sd, double delay, int64_t *seed, double t)
{
    const int32_t gtype = s->n->server_type; /* Parameters::server_type switches every class-Server object */
    if (type == 0) return t;                /* ConstServer, B/server.h:74 */
    if (type == 2) return t + mu + delay;   /* DetServer,   B/server.h:98 */
    if (type == 4) return t + loglogisticrandom(seed, mu, sd);   /* LogLogisticDelayServer, B/server.cpp:139-142 */
    if (gtype == 3) return t + (mu + lnrandom(seed, delay, sd)); /* Server::next, B/server.cpp:34-40 */
    if (gtype == 4) return t + loglogisticrandom(seed, mu, sd);  /* B/server.cpp:42-45 */
    double x = nrandom(seed, mu, sd);       /* Server::next, B/server.cpp:46-48 */
    return t + (x > 0.1 ? x : 0.1);         /* _MAX(min_hdway, x) = std::max(0.1, x) */
}

/* ChangeRateAction::execute (B/server.cpp:161-168) for every change of server sv whose time has come: the actions are booked
 * by readserverrates (B/network.cpp:5661), i.e. before every other event, so a change precedes the events of its own time
 * instant and equal-time changes run in file order. The reference mutates the Server; this restatement folds on demand. */
static void server_rates_at(const MzNet *n, int32_t sv, double t, double *mu, double *sd)
{
    *mu = n->srv_mu[sv];
    *sd = n->srv_sd[sv];
    double last = -1.0;
    for (;;) { /* next change of this server in (time, file order) after the ones already applied */
        int32_t best = -1;
        for (int32_t c = 0; c < n->n_rate_changes; ++c) {
            if (n->chg_server[c] != sv || !(n->chg_time[c] <= t)) continue;
            if (n->chg_time[c] < last) continue;
            if (best < 0 || n->chg_time[c] < n->chg_time[best]) best = c;
        }
        if (best < 0) break;
        /* apply every change at exactly that time in file order, then move past it */
        const double tt = n->chg_time[best];
        for (int32_t c = 0; c < n->n_rate_changes; ++c)
            if (n->chg_server[c] == sv && n->chg_time[c] == tt) {
                if (n->chg_mu[c] > 0.0) *mu = n->chg_mu[c];
                if (n->chg_sd[c] > 0.0) *sd = n->chg_sd[c];
            }
        last = nextafter(tt, INFINITY);
    }
}

/* ---------------- vehicles ---------------- */
static int32_t veh_nextlink(const OSim *s, int32_t v) /* Vehicle::nextlink, B/vehicle.cpp:52-58 */
{
    const OVeh *ve = &s->veh[v];
    const int32_t b = s->n->route_off[ve->route], e = s->n->route_off[ve->route + 1];
    const int32_t i = b + ve->pos + 1; /* pos = -1 while not entered → firstlink */
    return i < e ? s->n->route_links[i] : -1;
}

/* ---------------- Q ---------------- */
static void q_insert(OLink *l, int32_t at, QEnt e)
{
    if (l->size == l->cap) {
        l->cap = l->cap ? 2 * l->cap : 8;
        l->q = (QEnt *)realloc(l->q, sizeof(QEnt) * (size_t)l->cap);
    }
    memmove(l->q + at + 1, l->q + at, sizeof(QEnt) * (size_t)(l->size - at));
    l->q[at] = e;
    l->size++;
}
static void q_erase(OLink *l, int32_t at)
{
    memmove(l->q + at, l->q + at + 1, sizeof(QEnt) * (size_t)(l->size - at - 1));
    l->size--;
}
/* Q::enter_veh, B/q.cpp:52-72 */
static int q_enter(OLink *l, int32_t v, double ttime)
{
    if (l->size >= l->maxcap) return 0;
    QEnt e = { ttime, v };
    if (l->size == 0) { q_insert(l, 0, e); return 1; }
    int32_t i = l->size - 1;
    while (l->q[i].exit > ttime && i != 0) --i;
    q_insert(l, i + 1, e); /* insert(++viter): even when the head itself is later */
    return 1;
}
static int32_t q_nr_running(const OLink *l, double t) /* B/q.h:67-68 */
{
    int32_t i = 0;
    while (i < l->size && !(l->q[i].exit > t)) ++i;
    return l->size - i;
}
static double q_calc_space(const OSim *s, const OLink *l, double t) /* B/q.cpp:346-361 */
{
    double space = 0.0;
    for (int32_t i = l->size - 1; i >= 0; --i) {
        if (l->q[i].exit > t) return space;
        space += s->veh[l->q[i].veh].length;
    }
    return space;
}

/* ---------------- Link ---------------- */
static int link_full_t(OLink *l, double t) /* Link::full(time), B/link.cpp:155-183 */
{
    if (l->blocked_until == -2.0) return 1;
    if (l->size >= l->maxcap) {
        if (l->nr_exits_blocked > 0) l->blocked_until = -2.0;
        return 1;
    }
    if (l->blocked_until == -1.0) return 0;
    if (l->blocked_until <= t) { l->blocked_until = -1.0; return 0; }
    return 1;
}
static double link_density(const MzNet *n, const OLink *l, int32_t li) /* Link::density, B/link.cpp:657-661 */
{
    return l->size / (n->link_lanes[li] * (n->link_length[li] / 1000.0));
}
static double density_running_only(const OSim *s, int32_t li, double t) /* B/link.cpp:670-687 */
{
    const OLink *l = &s->lk[li];
    int32_t nr = q_nr_running(l, t);
    double queue_space = q_calc_space(s, l, t) / s->n->link_lanes[li];
    double running_length = s->n->link_length[li] - queue_space;
    if (nr && running_length) return nr / (running_length * (s->n->link_lanes[li] / 1000.0));
    return 0;
}
/* Bustrip::book_stop_visit + the host's transit model in one: the visit is logged, and the bus leaves the stop `dwell` later
   (Busstop::execute's enter branch computes that exit time and books the exit event, B/busline.cpp:1277-1298) */
static void heap_push(OSim *s, double t, int32_t act);
static void bus_book_visit(OSim *s, int32_t v, int32_t k, int32_t li, double t_enter, double t_arrive)
{
    mz_stop_visit r = { v, k, li, 0, t_enter, t_arrive };
    s->visits[s->n_visits++] = r;
    s->st.n_stop_visits++;
    const double dw = s->dwell ? s->dwell[s->n->trip_stop_off[v] + k] : 0.0;
    heap_push(s, t_arrive + dw, s->bus_act[v]);
}
static int link_enter_veh(OSim *s, int32_t li, int32_t v, double t) /* Link::enter_veh, B/link.cpp:418-523 */
{
    double ro = density_running_only(s, li, t);
    double speed = sd_speed(s->n, s->lk[li].sd_cur, ro);
    double exit_time = t + (s->n->link_length[li] / speed);
    OVeh *ve = &s->veh[v];
    ve->exit = exit_time;
    ve->pos += 1; /* set_curr_link(this): the link just entered is the next one on the route */
    ve->entry = t;
    if (s->n->trip_stop_off) { /* type-4 vehicle whose next stop is on this link: book the visit, do not queue (B/link.cpp:489-516) */
        const int32_t s0 = s->n->trip_stop_off[v], s1 = s->n->trip_stop_off[v + 1];
        const int32_t k = s0 < s1 ? s->bus_next[v] : 0;
        if (s0 + k < s1 && s->n->stop_link[s0 + k] == li) {
            double stop_position = s->n->stop_pos[s0 + k];
            double time_to_stop = t + ((exit_time - t) * (stop_position / s->n->link_length[li]));
            s->bus_next[v] = k + 1;
            bus_book_visit(s, v, k, li, t, time_to_stop);
            s->st.n_traversals++;
            return 1;
        }
    }
    int ok = q_enter(&s->lk[li], v, exit_time);
    if (ok) s->st.n_traversals++;
    return ok;
}
/* Busstop::execute, the bus leaves its stop (B/busline.cpp:1196-1263) */
static void bus_depart(OSim *s, int32_t v, double time)
{
    const MzNet *n = s->n;
    OVeh *ve = &s->veh[v];
    const int32_t li = n->route_links[n->route_off[ve->route] + ve->pos];
    const int32_t s0 = n->trip_stop_off[v], s1 = n->trip_stop_off[v + 1], k = s->bus_next[v];
    double position = n->stop_pos[s0 + k - 1];
    double ro = density_running_only(s, li, time);
    double speed = sd_speed(n, s->lk[li].sd_cur, ro);
    double link_total_travel_time = n->link_length[li] / speed;
    if (s0 + k < s1 && n->stop_link[s0 + k] == li) { /* the next stop IS on this link */
        double relative_length = (n->stop_pos[s0 + k] - position) / n->link_length[li];
        s->bus_next[v] = k + 1;
        bus_book_visit(s, v, k, li, ve->entry, time + link_total_travel_time * relative_length);
    } else {
        double relative_length = (n->link_length[li] - position) / n->link_length[li];
        double exit_time = time + link_total_travel_time * relative_length;
        ve->exit = exit_time;
        if (!q_enter(&s->lk[li], v, exit_time)) s->st.n_bus_refused++; /* Q::enter_veh returns false: the bus is lost */
    }
}
/* bookkeeping shared by both Link::exit_veh variants, B/link.cpp:536-581 / 601-648 */
static void link_exit_bookkeeping(OSim *s, int32_t li, int32_t v, double t)
{
    OLink *l = &s->lk[li];
    OVeh *ve = &s->veh[v];
    const int32_t P = s->n->n_periods;
    double entrytime = ve->entry, traveltime = t - entrytime;
    l->avg_time = (l->nr_passed * l->avg_time + traveltime) / (l->nr_passed + 1);
    if ((l->curr_period + 1) * s->n->periodlength < entrytime) {
        if (l->tmp_avg == 0.0)
            l->tmp_avg = l->curr_period < P ? s->n->link_hist[(size_t)li * P + l->curr_period] : 0.0;
        if (l->curr_period < P) s->avgtimes[(size_t)li * P + l->curr_period] = l->tmp_avg;
        l->curr_period++;
        l->tmp_avg = 0.0;
        l->tmp_passed = 0;
    } else {
        l->tmp_avg = (l->tmp_passed * l->tmp_avg + traveltime) / (l->tmp_passed + 1);
        l->tmp_passed++;
    }
    l->nr_passed++;
    ve->meters += s->n->link_length[li];
    if (s->log_n == s->log_cap) {
        s->log_cap = s->log_cap ? 2 * s->log_cap : 4096;
        s->log = (MzExit *)realloc(s->log, sizeof(MzExit) * (size_t)s->log_cap);
    }
    MzExit r = { v + 1, li, entrytime, t };
    s->log[s->log_n++] = r;
    s->st.n_exits++;
}
/* Link::exit_veh(time,next,lookback) → Q::exit_veh, B/link.cpp:528-589, B/q.cpp:156-222. Returns veh or -1. */
static int32_t link_exit_veh_turn(OSim *s, int32_t li, double t, int32_t next, int32_t lookback)
{
    OLink *l = &s->lk[li];
    l->q_ok = 0;
    l->q_next_action = -1.0;
    if (l->size == 0) return -1;
    int32_t n = 0, i = 0, found = 0;
    while (i < l->size) {
        if (veh_nextlink(s, l->q[i].veh) == next) { found = 1; break; }
        ++n; ++i;
    }
    if (!found) { l->q_next_action = t + l->freeflow; return -1; }
    if (l->q[i].exit <= t) {
        if (n < lookback) {
            int32_t v = l->q[i].veh;
            q_erase(l, i);
            l->q_ok = 1;
            link_exit_bookkeeping(s, li, v, t);
            return v;
        }
        /* next_action = time + 1.0*(n / nextlink->nr_lanes) — integer division, B/q.cpp:204 */
        l->q_next_action = t + 1.0 * (n / s->n->link_lanes[next]);
        return -1;
    }
    l->q_next_action = l->q[i].exit;
    return -1;
}
static int32_t link_exit_veh_head(OSim *s, int32_t li, double t) /* B/link.cpp:593-655, B/q.cpp:227-249 */
{
    OLink *l = &s->lk[li];
    l->q_ok = 0;
    if (l->size == 0) return -1;
    l->q_next_action = 0.0;
    if (l->q[0].exit <= t) {
        int32_t v = l->q[0].veh;
        q_erase(l, 0);
        l->q_ok = 1;
        link_exit_bookkeeping(s, li, v, t);
        return v;
    }
    l->q_next_action = l->q[0].exit;
    return -1;
}
static double link_next_action(const OLink *l, double t) /* B/link.cpp:188-197 */
{
    return l->size == 0 ? t + l->freeflow : l->q_next_action;
}
/* Link::update_exit_times + Q::update_exit_times, B/link.cpp:320-390, B/q.cpp:74-129 */
static void link_update_exit_times(OSim *s, int32_t li, double t, int32_t next, int32_t lookback)
{
    const MzNet *n = s->n;
    OLink *l = &s->lk[li];
    const double kjam = 142.0;
    int32_t f = l->sd_cur;
    double k_dn = link_density(n, &s->lk[next], next);
    double v_exit = sd_speed(n, f, kjam);
    double v_dn = sd_speed(n, s->lk[next].sd_cur, k_dn);
    double v_sw;
    if (v_dn <= v_exit) v_dn = v_exit - 1.0;
    if (k_dn >= n->sd_romax[f]) v_sw = v_exit - 1.0;
    else v_sw = (k_dn * v_dn) / (kjam - k_dn);
    if (v_sw >= v_exit) v_sw = v_exit - 1.0;
    if (v_sw == 0.0) v_sw = v_exit;
    /* Q::update_exit_times */
    {
        int32_t cnt = 0, i = 0;
        double t0 = t, t1, t2, newtime;
        while (i < l->size) {
            if (l->q[i].exit > t0) break;
            int32_t v = l->q[i].veh;
            if (cnt < lookback) {
                cnt++;
                if (veh_nextlink(s, v) == next) {
                    t1 = s->veh[v].length / v_sw;
                    t2 = s->veh[v].length / v_exit;
                    newtime = t0 + t1 + t2;
                    l->q[i].exit = newtime;
                    s->veh[v].exit = newtime;
                    t0 = newtime;
                }
                i++;
            } else {
                cnt++;
                t1 = s->veh[v].length / v_sw;
                t2 = s->veh[v].length / v_exit;
                newtime = t0 + t1 + t2;
                int32_t lanes = n->link_lanes[li]; /* vehicle->get_curr_link()->get_nr_lanes() */
                for (int32_t k = 0; k < lanes && i < l->size; ++k) {
                    l->q[i].exit = newtime;
                    s->veh[l->q[i].veh].exit = newtime;
                    i++;
                }
                t0 = newtime;
            }
        }
    }
    if (l->blocked_until == -2.0) l->blocked_until = (n->link_length[li] / v_sw) + t;
}

/* ---------------- actions ---------------- */
static double turn_execute(OSim *s, int32_t k, double t) /* TurnAction::execute + Turning::process_veh */
{
    const MzNet *n = s->n;
    const int32_t in = n->turn_in[k], out = n->turn_out[k], sv = n->turn_server[k];
    OLink *lin = &s->lk[in];
    double nexttime = t;
    int ok = 0;
    if (link_full_t(&s->lk[out], t)) {
        if (!s->turn_blocked[k]) { s->turn_blocked[k] = 1; lin->nr_exits_blocked++; }
        nexttime = t + 1.0;
    } else {
        if (s->turn_blocked[k]) {
            s->turn_blocked[k] = 0;
            link_update_exit_times(s, in, t, out, n->turn_lookback[k]);
            lin->nr_exits_blocked--;
        }
        if (lin->size == 0) {
            nexttime = t + lin->freeflow;
        } else {
            int32_t v = link_exit_veh_turn(s, in, t, out, n->turn_lookback[k]);
            if (lin->q_ok) {
                double delay = n->srv_delay[sv];
                ok = link_enter_veh(s, out, v, t + delay);
                if (!ok) nexttime = link_next_action(lin, t); /* vehicle dropped; cannot happen after full()==false */
            } else {
                nexttime = link_next_action(lin, t);
            }
        }
    }
    double new_time;
    if (ok) {
        double mu, sd;
        server_rates_at(n, sv, t, &mu, &sd);
        new_time = server_next(s, n->srv_type[sv], mu, sd, n->srv_delay[sv], &s->srv_seed[sv], t);
    }
    else {
        new_time = nexttime;
        if (new_time < t) new_time = t + 0.1;
    }
    return new_time;
}

static double dest_execute(OSim *s, int32_t d, int32_t li, double t) /* Daction::execute, B/node.cpp:637-673 */
{
    const MzNet *n = s->n;
    int32_t v = link_exit_veh_head(s, li, t);
    if (s->lk[li].q_ok) {
        s->arrival[v] = t; /* Vehicle::report(time) */
        s->st.n_arrived++;
        int32_t sv = n->dest_server[d];
        if (sv >= 0) {
            double mu, sd;
            server_rates_at(n, sv, t, &mu, &sd);
            return server_next(s, n->srv_type[sv], mu, sd, n->srv_delay[sv], &s->srv_seed[sv], t);
        }
        /* Destination's own default: new Server(-20000, 0, 0.1, 0.5, 0.001) — class Server, i.e. N(0.1,0.5)>=0.1 */
        return server_next(s, 1, 0.1, 0.5, 0.001, &s->srv_seed[n->n_servers + d], t);
    }
    return link_next_action(&s->lk[li], t);
}

/* Origin::insert_veh / transfer_veh with the InputLink virtual queue, B/node.cpp:126-250 */
static void vq_push(OSim *s, int32_t o, int32_t v)
{
    if (s->vq_size[o] == s->vq_cap[o]) {
        s->vq_cap[o] = s->vq_cap[o] ? 2 * s->vq_cap[o] : 16;
        s->vq[o] = (int32_t *)realloc(s->vq[o], sizeof(int32_t) * (size_t)s->vq_cap[o]);
    }
    s->vq[o][s->vq_size[o]++] = v; /* exit_time = time, non-decreasing ⇒ Q::enter_veh appends */
}
static void origin_insert(OSim *s, int32_t v, double t)
{
    const MzNet *n = s->n;
    const int32_t o = n->trip_origin[v];
    const int32_t first = n->route_links[n->route_off[n->trip_route[v]]];
    OLink *l = &s->lk[first];
    s->veh[v].exit = t; /* only InputLink::enter_veh stamps these, harmless otherwise */
    if (l->size >= l->maxcap) { /* link->full() (no time argument): park in the virtual queue */
        s->veh[v].entry = t;
        vq_push(s, o, v);
        return;
    }
    if (s->vq_size[o] == 0) {
        link_enter_veh(s, first, v, t);
        return;
    }
    s->veh[v].entry = t;
    vq_push(s, o, v);
    /* transfer_veh(link,time): drain head-first every parked vehicle whose first link is `first` while not full */
    for (;;) {
        if (s->vq_size[o] == 0 || l->size >= l->maxcap) return;
        int32_t i = 0, found = -1;
        for (; i < s->vq_size[o]; ++i)
            if (n->route_links[n->route_off[n->trip_route[s->vq[o][i]]]] == first) { found = i; break; }
        if (found < 0) return; /* exit_ok false → stop */
        int32_t w = s->vq[o][found];
        memmove(s->vq[o] + found, s->vq[o] + found + 1, sizeof(int32_t) * (size_t)(s->vq_size[o] - found - 1));
        s->vq_size[o]--;
        link_enter_veh(s, first, w, t);
    }
}

#ifdef ORC_TRACE
#include <stdio.h>
static void orc_trace(double t, int kind, int idx) { static FILE *f; if (!f) f = fopen("/tmp/orc_trace.txt", "w"); fprintf(f, "%.17g %d %d\n", t, kind, idx); fflush(f); }
#endif
/* ---------------- public ---------------- */
void orc_sim_destroy(OSim *s)
{
    if (!s) return;
    for (int32_t i = 0; i < s->n->n_links; ++i) free(s->lk[i].q);
    for (int32_t i = 0; i < s->n->n_origins; ++i) free(s->vq[i]);
    free(s->lk); free(s->veh); free(s->srv_seed); free(s->turn_blocked); free(s->vq); free(s->vq_size);
    free(s->od_off); free(s->od_trips); free(s->od_next);
    free(s->turn_active); free(s->turn_act0); free(s->turn_nact); free(s->sigc_active); free(s->sigp_active);
    free(s->sigs_active); free(s->sigc_cur); free(s->sigp_cur); free(s->sigc_act); free(s->sigp_act); free(s->sigs_act);
    free(s->bus_next); free(s->bus_act); free(s->visits);
    free(s->vq_cap); free(s->acts); free(s->heap); free(s->avgtimes); free(s->arrival); free(s->log);
    free(s);
}

OSim *orc_sim_create(const MzNet *n)
{
    OSim *s = (OSim *)calloc(1, sizeof(OSim));
    s->n = n;
    s->lk = (OLink *)calloc((size_t)n->n_links + 1, sizeof(OLink));
    s->veh = (OVeh *)calloc((size_t)n->n_trips + 1, sizeof(OVeh));
    s->srv_seed = (int64_t *)calloc((size_t)n->n_servers + n->n_dests + 1, sizeof(int64_t));
    s->turn_blocked = (uint8_t *)calloc((size_t)n->n_turnings + 1, 1);
    s->turn_active = (uint8_t *)malloc((size_t)n->n_turnings + 1);
    memset(s->turn_active, 1, (size_t)n->n_turnings + 1); /* "turning is active by default", B/turning.cpp:22 ... */
    if (n->n_sig_stages > 0) /* ... but Stage::add_turning stops every turning a signal regulates (B/trafficsignal.h:75-76) */
        for (int32_t i = 0; i < n->sigs_turn_off[n->n_sig_stages]; ++i) s->turn_active[n->sigs_turns[i]] = 0;
    s->turn_act0 = (int32_t *)calloc((size_t)n->n_turnings + 1, sizeof(int32_t));
    s->turn_nact = (int32_t *)calloc((size_t)n->n_turnings + 1, sizeof(int32_t));
    s->sigc_active = (uint8_t *)calloc((size_t)n->n_sig_controls + 1, 1);
    s->sigp_active = (uint8_t *)calloc((size_t)n->n_sig_plans + 1, 1);
    s->sigs_active = (uint8_t *)calloc((size_t)n->n_sig_stages + 1, 1);
    s->sigc_cur = (int32_t *)calloc((size_t)n->n_sig_controls + 1, sizeof(int32_t));
    s->sigp_cur = (int32_t *)calloc((size_t)n->n_sig_plans + 1, sizeof(int32_t));
    s->sigc_act = (int32_t *)calloc((size_t)n->n_sig_controls + 1, sizeof(int32_t));
    s->sigp_act = (int32_t *)calloc((size_t)n->n_sig_plans + 1, sizeof(int32_t));
    s->sigs_act = (int32_t *)calloc((size_t)n->n_sig_stages + 1, sizeof(int32_t));
    s->vq = (int32_t **)calloc((size_t)n->n_origins + 1, sizeof(int32_t *));
    s->vq_size = (int32_t *)calloc((size_t)n->n_origins + 1, sizeof(int32_t));
    s->vq_cap = (int32_t *)calloc((size_t)n->n_origins + 1, sizeof(int32_t));
    s->avgtimes = (double *)malloc(sizeof(double) * ((size_t)n->n_links * n->n_periods + 1));
    s->arrival = (double *)malloc(sizeof(double) * ((size_t)n->n_trips + 1));
    memcpy(s->avgtimes, n->link_hist, sizeof(double) * (size_t)n->n_links * n->n_periods); /* new LinkTime(*histtimes) */
    for (int32_t i = 0; i < n->n_trips; ++i) {
        s->arrival[i] = -1.0;
        s->veh[i].route = n->trip_route[i];
        s->veh[i].pos = -1;
        s->veh[i].length = n->trip_length[i];
        s->veh[i].start = n->trip_time[i];
    }
    if (n->trip_stop_off) {
        s->bus_next = (int32_t *)calloc((size_t)n->n_trips + 1, sizeof(int32_t));
        s->bus_act = (int32_t *)calloc((size_t)n->n_trips + 1, sizeof(int32_t));
        s->visits = (mz_stop_visit *)calloc((size_t)n->trip_stop_off[n->n_trips] + 1, sizeof(mz_stop_visit));
    }
    for (int32_t i = 0; i < n->n_servers + n->n_dests; ++i) s->srv_seed[i] = n->randseed;
    for (int32_t i = 0; i < n->n_links; ++i) {
        OLink *l = &s->lk[i];
        l->maxcap = (int32_t)(n->link_length[i] * n->link_lanes[i] / n->standard_veh_length); /* int division */
        l->freeflow = n->link_length[i] / sd_speed(n, n->link_sdfunc[i], 0.0);
        if (l->freeflow < 1.0) l->freeflow = 1.0;
        l->blocked_until = -1.0;
        l->sd_cur = l->sd_tmp = n->link_sdfunc[i];
    }
    return s;
}

/* ---------------- traffic signals: B/trafficsignal.cpp + Turning::activate / stop (B/turning.cpp:218-227, B/turning.h:60) ----
 * A red turning's TurnActions die at their next event (TurnAction::execute does nothing and does not re-book, B/turning.cpp:
 * 236-262); Turning::activate executes every TurnAction on the spot, 3e-8 s apart, and each execution books the action again —
 * on top of whatever booking of the same action is still pending (a chain whose next event lies beyond the red phase
 * survives it, and the next green adds a second chain). Signal-regulated turnings start red (Stage::add_turning). */
static void turning_activate(OSim *s, int32_t k, double time) /* Turning::activate */
{
    s->turn_active[k] = 1;
    for (int32_t j = 0; j < s->turn_nact[k]; ++j) {
        heap_push(s, turn_execute(s, k, time), s->turn_act0[k] + j);
        time += 0.00000003;
    }
}

static void stage_stop(OSim *s, int32_t st) /* Stage::stop */
{
    const MzNet *n = s->n;
    s->sigs_active[st] = 0;
    for (int32_t i = n->sigs_turn_off[st]; i < n->sigs_turn_off[st + 1]; ++i) s->turn_active[n->sigs_turns[i]] = 0;
}

static void stage_execute(OSim *s, int32_t st, double time) /* Stage::execute */
{
    const MzNet *n = s->n;
    if (!s->sigs_active[st]) {
        for (int32_t i = n->sigs_turn_off[st]; i < n->sigs_turn_off[st + 1]; ++i) turning_activate(s, n->sigs_turns[i], time);
        s->sigs_active[st] = 1;
        heap_push(s, time + n->sigs_duration[st], s->sigs_act[st]); /* the event that stops this stage */
    } else {
        stage_stop(s, st);
    }
}

static void plan_execute(OSim *s, int32_t p, double time) /* SignalPlan::execute */
{
    const MzNet *n = s->n;
    const int32_t s0 = n->sigp_stage_off[p], s1 = n->sigp_stage_off[p + 1];
    if (!s->sigp_active[p]) {
        s->sigp_cur[p] = s0; /* start at the first stage when this plan becomes active */
        s->sigp_active[p] = 1;
    }
    if (time < n->sigp_stop[p]) {
        const int32_t cur = s->sigp_cur[p];
        stage_execute(s, cur, time);
        const double next_stage_time = time + n->sigs_duration[cur];
        s->sigp_cur[p] = cur + 1 == s1 ? s0 : cur + 1; /* cycle round */
        heap_push(s, next_stage_time, s->sigp_act[p]);
    } else { /* at end of signal plan turn all stages to red */
        s->sigp_active[p] = 0;
        for (int32_t st = s0; st < s1; ++st) stage_stop(s, st);
    }
}

static void control_execute(OSim *s, int32_t c, double time) /* SignalControl::execute */
{
    const MzNet *n = s->n;
    if (!s->sigc_active[c]) {
        s->sigc_cur[c] = n->sigc_plan_off[c];
        s->sigc_active[c] = 1;
    }
    plan_execute(s, s->sigc_cur[c], time);
    s->sigc_cur[c]++;
    if (s->sigc_cur[c] < n->sigc_plan_off[c + 1]) heap_push(s, n->sigp_start[s->sigc_cur[c]], s->sigc_act[c]);
}

/* Network::init (first events) + Network::step (run loop). Returns 0. */
int orc_sim_run(OSim *s)
{
    const MzNet *n = s->n;
    /* actions: turnactions (min(lanes_in,lanes_out) per turning), dactions (one per lane per incoming link) */
    int32_t na = 1 + n->n_odpairs + n->n_sig_controls + n->n_sig_plans + n->n_sig_stages + n->n_incidents;
    for (int32_t k = 0; k < n->n_turnings; ++k) {
        int32_t a = n->link_lanes[n->turn_in[k]], b = n->link_lanes[n->turn_out[k]];
        na += a < b ? a : b;
    }
    for (int32_t d = 0; d < n->n_dests; ++d)
        for (int32_t l = 0; l < n->n_links; ++l)
            if (n->link_out_node[l] == n->dest_node[d]) na += n->link_lanes[l];
    if (n->trip_stop_off)
        for (int32_t v = 0; v < n->n_trips; ++v) na += n->trip_stop_off[v] < n->trip_stop_off[v + 1];
    s->acts = (Act *)malloc(sizeof(Act) * (size_t)na);
    s->n_acts = 0;
    if (n->trip_stop_off) /* one departure action per bus; it has events only while the bus stands at a stop */
        for (int32_t v = 0; v < n->n_trips; ++v)
            if (n->trip_stop_off[v] < n->trip_stop_off[v + 1]) {
                s->bus_act[v] = s->n_acts;
                s->acts[s->n_acts].kind = ACT_BUSDEP; s->acts[s->n_acts++].idx = v;
            }
    double initvalue = 0.1; /* B/network.cpp:7660 */
    for (int32_t k = 0; k < n->n_turnings; ++k) { /* Turning::init: each action executes NOW, 3e-8 apart */
        int32_t a = n->link_lanes[n->turn_in[k]], b = n->link_lanes[n->turn_out[k]];
        int32_t cnt = a < b ? a : b;
        double t = initvalue;
        s->turn_act0[k] = s->n_acts;
        s->turn_nact[k] = cnt;
        for (int32_t j = 0; j < cnt; ++j) {
            int32_t id = s->n_acts++;
            s->acts[id].kind = ACT_TURN; s->acts[id].idx = k;
            if (s->turn_active[k]) { /* a red turning's actions do nothing at init and are not booked (B/turning.cpp:236-262) */
                s->st.n_events++;
                g_tp = t; g_tpp = 0.0;
                heap_push(s, turn_execute(s, k, t), id);
            }
            t += 0.00000003;
        }
        initvalue += 0.00001;
    }
    /* signal controls: SignalControl::execute right here, each 0.000001 after the other (B/network.cpp:7666-7670) */
    for (int32_t st = 0; st < n->n_sig_stages; ++st) { s->sigs_act[st] = s->n_acts; s->acts[s->n_acts].kind = ACT_SIGS; s->acts[s->n_acts++].idx = st; }
    for (int32_t p = 0; p < n->n_sig_plans; ++p) { s->sigp_act[p] = s->n_acts; s->acts[s->n_acts].kind = ACT_SIGP; s->acts[s->n_acts++].idx = p; }
    for (int32_t c = 0; c < n->n_sig_controls; ++c) { s->sigc_act[c] = s->n_acts; s->acts[s->n_acts].kind = ACT_SIGC; s->acts[s->n_acts++].idx = c; }
    for (int32_t c = 0; c < n->n_sig_controls; ++c) {
        s->st.n_events++;
        g_tp = initvalue; g_tpp = 0.0;
        control_execute(s, c, initvalue);
        initvalue += 0.000001;
    }
    for (int32_t i = 0; i < n->n_transit_init_steps; ++i) initvalue += 0.00001; /* buslines, ODstops: B/network.cpp:7672-7694 */
    /* ODpair::execute books the pair's first ODaction here, in OD order (B/network.cpp:7702-7723, B/od.cpp:355-368);
       every OD pair is an action of its own, so its events take part in the FIFO order like everybody else's */
    s->od_off = (int32_t *)calloc((size_t)n->n_odpairs + 2, sizeof(int32_t));
    s->od_trips = (int32_t *)malloc(sizeof(int32_t) * ((size_t)n->n_trips + 1));
    s->od_next = (int32_t *)calloc((size_t)n->n_odpairs + 1, sizeof(int32_t));
    for (int32_t v = 0; v < n->n_trips; ++v) s->od_off[n->trip_od[v] + 1]++;
    for (int32_t k = 0; k < n->n_odpairs; ++k) s->od_off[k + 1] += s->od_off[k];
    {
        int32_t *c = (int32_t *)calloc((size_t)n->n_odpairs + 1, sizeof(int32_t));
        for (int32_t v = 0; v < n->n_trips; ++v) s->od_trips[s->od_off[n->trip_od[v]] + c[n->trip_od[v]]++] = v;
        free(c);
    }
    for (int32_t k = 0; k < n->n_odpairs; ++k) {
        int32_t id = s->n_acts++;
        s->acts[id].kind = ACT_TRIPS; s->acts[id].idx = k; s->acts[id].aux = 0;
        g_tp = initvalue; g_tpp = 0.0;
        if (s->od_off[k] < s->od_off[k + 1]) heap_push(s, n->trip_time[s->od_trips[s->od_off[k]]], id);
        initvalue += 0.00001;
    }
    /* destinations: Destination::execute runs every Daction at initvalue. dactions were built with insert(begin())
       over ascending link ids, one per lane ⇒ vector order = descending link id (B/node.cpp:318-332) */
    for (int32_t d = 0; d < n->n_dests; ++d) {
        for (int32_t l = n->n_links - 1; l >= 0; --l) {
            if (n->link_out_node[l] != n->dest_node[d]) continue;
            for (int32_t j = 0; j < n->link_lanes[l]; ++j) {
                int32_t id = s->n_acts++;
                s->acts[id].kind = ACT_DEST; s->acts[id].idx = l; s->acts[id].aux = d;
                s->st.n_events++;
                g_tp = initvalue; g_tpp = 0.0;
                heap_push(s, dest_execute(s, d, l, initvalue), id);
            }
        }
        initvalue += 0.00001;
    }
    /* Incident::Incident books its start while the incident file is read — AFTER Network::init (executemaster,
       B/network.cpp:7380-7383), so it follows every booking of init and precedes every booking of the run */
    for (int32_t i = 0; i < n->n_incidents; ++i) {
        int32_t id = s->n_acts++;
        s->acts[id].kind = ACT_INC; s->acts[id].idx = i;
        g_tp = initvalue; g_tpp = 0.0;
        heap_push(s, n->inc_start[i], id);
    }
    /* the big loop: while (time < runtime) time = eventlist->next();   B/network.cpp:7804-7806 */
    double time = 0.0;
    while (time < n->runtime && s->heap_n > 0) {
        /* a MatrixAction or the poll of an inactive ODaction (neither is an action of this port: the host pre-pass folds
           them into the trip table) may be the event that ends the loop: MzNet::phantom_final_t */
        if (n->phantom_final_t > 0.0) {
            const Ev top = s->heap[0];
            if (!(top.t < n->runtime) &&
                (n->phantom_final_t < top.t || (n->phantom_final_t == top.t && n->phantom_final_tp < top.tp))) {
                time = n->phantom_final_t;
                break;
            }
        }
        Ev e = heap_pop(s);
        time = e.t;
        Act a = s->acts[e.act];
        s->st.n_events++;
        g_tp = e.t; g_tpp = e.tp;
#ifdef ORC_TRACE
        if (a.kind != ACT_TRIPS) orc_trace(time, a.kind, a.idx);
#endif
        if (a.kind == ACT_TURN) {
            /* TurnAction::execute: a red turning's action does nothing and is not booked again (B/turning.cpp:236-262) */
            if (s->turn_active[a.idx]) heap_push(s, turn_execute(s, a.idx, time), e.act);
        } else if (a.kind == ACT_SIGC) {
            control_execute(s, a.idx, time);
        } else if (a.kind == ACT_SIGP) {
            plan_execute(s, a.idx, time);
        } else if (a.kind == ACT_SIGS) {
            stage_execute(s, a.idx, time);
        } else if (a.kind == ACT_DEST) {
            heap_push(s, dest_execute(s, a.aux, a.idx, time), e.act);
        } else if (a.kind == ACT_BUSDEP) {
            bus_depart(s, a.idx, time);
        } else if (a.kind == ACT_INC) { /* Incident::execute without information broadcast, B/network.cpp:8092-8112 */
            OLink *l = &s->lk[n->inc_link[a.idx]];
            if (time < n->inc_stop[a.idx]) { /* Link::set_incident: the member blocked_until is NOT touched (the parameter shadows it) */
                l->sd_tmp = l->sd_cur;
                l->sd_cur = n->inc_sdfunc[a.idx];
                heap_push(s, n->inc_stop[a.idx], e.act);
            } else { /* Link::unset_incident */
                l->sd_cur = l->sd_tmp;
                l->blocked_until = -1.0;
            }
        } else {
            const int32_t k = a.idx;
            const int32_t v = s->od_trips[s->od_off[k] + s->od_next[k]];
            origin_insert(s, v, time);
            s->od_next[k]++;
            if (s->od_off[k] + s->od_next[k] < s->od_off[k + 1])
                heap_push(s, n->trip_time[s->od_trips[s->od_off[k] + s->od_next[k]]], e.act);
        }
    }
    s->st.end_time = time;
    return 0;
}

void orc_sim_set_dwell(OSim *s, const double *dwell) { s->dwell = dwell; }
int64_t orc_sim_stop_visits(const OSim *s, const mz_stop_visit **out)
{
    *out = s->visits;
    return s->n_visits;
}
const mz_sim_stats *orc_sim_stats(const OSim *s) { return &s->st; }
const double *orc_sim_arrivals(const OSim *s) { return s->arrival; }
const double *orc_sim_linktimes(const OSim *s) { return s->avgtimes; }
int64_t orc_sim_exit_log(const OSim *s, const MzExit **out)
{
    *out = s->log;
    return s->log_n;
}
