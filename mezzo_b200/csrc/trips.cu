// Row f1 of SURVEY.md §8: the host pre-pass that replaces the ODaction event chain of the reference — trip generation and
// route choice are independent of the traffic state, so the whole trip table of MzNet is produced before the simulation
// starts. Plain host C++ (no device work), native because the north-star demand is 10 M trips.
//
// Reference (B/ = /root/reference/branches/busmezzo/mezzo_lib/src/):
//   ODpair::execute (first booking)         B/od.cpp:355-368     per-pair Random seeded 42 (_DETERMINISTIC_ROUTE_CHOICE, B/od.cpp:313-320)
//   ODaction::execute / ODServer::next      B/od.cpp:69-108, B/server.cpp:82-91 (own Random seeded with randseed, B/server.cpp:6-13)
//   ODpair::select_route                    B/od.cpp:247-300     Kirchhoff shares pow(cost, kirchoff_alpha), one urandom draw
//   Route::cost + LinkTime::cost            B/route.cpp:173-197 (inverted cache test), B/linktimes.cpp:36-82
//   Random::urandom / rrandom               B/Random.cpp:85-98, 284
//   Network::init order of the init times   B/network.cpp:7657-7731 (0.1 + 0.00001 per Turning::init, then per ODpair)
#include <algorithm>
#include <cmath>
#include <numeric>
#include <vector>

#include "../../include/mezzo_b200.h"
#include "common.h"

namespace {

struct Lcg {  // Random::urandom, B/Random.cpp:85-98 (Park-Miller, A = 48271, M = 2^31 - 1)
    long long seed;
    double urandom()
    {
        const long long M = 2147483647LL, A = 48271LL, Q = M / A, R = M % A;
        seed = A * (seed % Q) - R * (seed / Q);
        if (seed <= 0) seed += M;
        return (double)seed / (double)M;
    }
};

// Route::cost (B/route.cpp:173-197): recomputed on every call until a call comes more than update_interval_routes after the
// last computation — from then on the stale sum is returned for good (the cache test is inverted in the reference)
struct RouteCost {
    const int32_t *links;
    int n;
    double sumcost = 0.0, last = 0.0;
    double cost(const MzDemand *d, double time)
    {
        if (sumcost > 0.0 && last + d->update_interval_routes < time) return sumcost;
        double temp = time;
        last = time;
        for (int i = 0; i < n; ++i) {
            const int p = (int)(temp / d->periodlength);  // static_cast<int>; times are never negative
            temp += (p >= 0 && p < d->n_periods) ? d->link_hist[(size_t)links[i] * d->n_periods + p] : 0.1;
        }
        sumcost = temp - time;
        return sumcost;
    }
};

}  // namespace

extern "C" int mz_trips_generate(const MzDemand *d, int64_t cap, double *time, double *prev_time, int32_t *od,
                                 int32_t *route, int64_t *n_out, int32_t *n_odpairs_initialised)
{
    MZ_REQUIRE(d && n_out && n_odpairs_initialised, MZ_EINVAL, "null argument");
    MZ_REQUIRE(d->n_odpairs >= 0 && d->n_periods >= 1 && d->periodlength > 0, MZ_EINVAL, "bad demand header");
    double initvalue = 0.1;
    for (int i = 0; i < d->n_turnings; ++i) initvalue += 0.00001;  // one per Turning::init call, accumulated like the reference
    for (int i = 0; i < d->n_signal_controls; ++i) initvalue += 0.000001;  // one per SignalControl (B/network.cpp:7666-7670)
    std::vector<double> t_all, tp_all;
    std::vector<int32_t> od_all, route_all;
    int32_t n_init = 0;
    std::vector<RouteCost> rc;
    std::vector<double> util;
    for (int k = 0; k < d->n_odpairs; ++k) {
        const int r0 = d->od_route_off[k], r1 = d->od_route_off[k + 1];
        if (r1 == r0) continue;  // "chuck out the OD pairs without paths": no init slot is consumed
        Lcg rnd{42}, srv{(long long)d->randseed};
        const double t_init = initvalue;
        initvalue += 0.00001;
        ++n_init;
        const double rate = d->od_rate[k];
        if (rate <= 0) continue;
        const double temp = 3600 / rate;
        double t = t_init + (1.2 + (temp - 1.2) * rnd.urandom());  // urandom(1.2, 3600/rate)
        if (rate < 1) continue;  // ODaction inactive: it only re-polls, never generates
        const double mu = 3600.0 / rate;
        rc.clear();
        for (int r = r0; r < r1; ++r) {
            const int ri = d->od_routes[r];
            RouteCost c;
            c.links = d->route_links + d->route_off[ri];
            c.n = d->route_off[ri + 1] - d->route_off[ri];
            // cache state the Route object carries into the run (ODpair::delete_spurious_routes has called cost() before)
            if (d->route_sumcost0) c.sumcost = d->route_sumcost0[ri];
            if (d->route_last0) c.last = d->route_last0[ri];
            rc.push_back(c);
        }
        util.resize(rc.size());
        double prev = t_init;  // the booking event: ODpair::execute, then the previous ODaction
        for (;;) {
            // ODaction::execute → select_route(time): no draw with a single route; otherwise one draw of the pair's own Random
            int choice_idx = 0;
            if (rc.size() > 1) {
                double total = 0.0;
                for (size_t i = 0; i < rc.size(); ++i) {
                    util[i] = std::pow(rc[i].cost(d, t), d->kirchoff_alpha);
                    total += util[i];
                }
                const double choice = rnd.urandom();
                double sub = 0.0;
                choice_idx = 0;  // "TROUBLE: No Route selected! Returning the first in list"
                for (size_t i = 0; i < rc.size(); ++i) {
                    sub += util[i];
                    if (sub / total >= choice) {
                        choice_idx = (int)i;
                        break;
                    }
                }
            }
            t_all.push_back(t);
            tp_all.push_back(prev);
            od_all.push_back(k);
            route_all.push_back(d->od_routes[r0 + choice_idx]);
            if (t >= d->runtime) break;  // the first event at/after the horizon may still execute (B/network.cpp:7804-7806)
            prev = t;
            // ODServer::next (B/server.cpp:82-87): time + mu, or time + rrandom(mu) = -log(urandom) * mu (B/Random.cpp:284)
            t = d->od_servers_deterministic ? t + mu : t + (-std::log(srv.urandom()) * mu);
        }
    }
    *n_odpairs_initialised = n_init;
    const int64_t n = (int64_t)t_all.size();
    *n_out = n;
    if (cap == 0 && !time) return MZ_OK;  // sizing call
    MZ_REQUIRE(time && prev_time && od && route, MZ_EINVAL, "null output buffer");
    MZ_REQUIRE(cap >= n, MZ_ELIMIT, "%lld trips, buffers hold %lld", (long long)n, (long long)cap);
    // global event order: by time; equal times FIFO = by the booking event's time, then OD init order (stable)
    std::vector<int64_t> order((size_t)n);
    std::iota(order.begin(), order.end(), (int64_t)0);
    std::stable_sort(order.begin(), order.end(), [&](int64_t a, int64_t b) {
        if (t_all[a] != t_all[b]) return t_all[a] < t_all[b];
        return tp_all[a] < tp_all[b];
    });
    for (int64_t i = 0; i < n; ++i) {
        const int64_t s = order[i];
        time[i] = t_all[s];
        prev_time[i] = tp_all[s];
        od[i] = od_all[s];
        route[i] = route_all[s];
    }
    return MZ_OK;
}
