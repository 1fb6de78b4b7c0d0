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
//   demand slices: MatrixAction::execute    B/network.cpp:8213-8229 → ODpair::set_rate / ODaction::set_rate (B/od.h:63,102);
//   inactive ODactions (rate < 1) poll      B/od.cpp:102-107
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
    for (int i = 0; i < d->n_transit_init_steps; ++i) initvalue += 0.00001;  // buslines, active ODstops (B/network.cpp:7672-7694)
    std::vector<double> t_all, tp_all;
    std::vector<int32_t> od_all, route_all;
    int32_t n_init = 0;
    std::vector<RouteCost> rc;
    std::vector<double> util;
    // demand slices per OD pair (execution order is the input order); the MatrixActions themselves are events too
    std::vector<int> sl_off((size_t)d->n_odpairs + 1, 0), sl_idx((size_t)std::max(0, d->n_slice_rates));
    double ph_t = 0.0, ph_tp = 0.0;
    bool have_ph = false;
    auto phantom = [&](double t, double tp) {
        if (!have_ph || t < ph_t || (t == ph_t && tp < ph_tp)) {
            ph_t = t;
            ph_tp = tp;
            have_ph = true;
        }
    };
    for (int i = 0; i < d->n_slice_rates; ++i) {
        MZ_REQUIRE(d->slice_od && d->slice_time && d->slice_rate, MZ_EINVAL, "slice arrays missing");
        MZ_REQUIRE(d->slice_od[i] >= 0 && d->slice_od[i] < d->n_odpairs, MZ_EINVAL, "slice entry %d: OD pair out of range", i);
        MZ_REQUIRE(d->slice_rate[i] >= 0, MZ_EINVAL, "slice entry %d: negative rate", i);
        MZ_REQUIRE(i == 0 || d->slice_time[i] >= d->slice_time[i - 1], MZ_EINVAL, "slice entries must be in time order");
        sl_off[(size_t)d->slice_od[i] + 1]++;
        if (d->slice_time[i] >= d->runtime) phantom(d->slice_time[i], -INFINITY);
    }
    for (int k = 0; k < d->n_odpairs; ++k) sl_off[(size_t)k + 1] += sl_off[(size_t)k];
    {
        std::vector<int> c(sl_off.begin(), sl_off.end() - 1);
        for (int i = 0; i < d->n_slice_rates; ++i) sl_idx[(size_t)c[(size_t)d->slice_od[i]]++] = i;
    }
    for (int k = 0; k < d->n_odpairs; ++k) {
        const int r0 = d->od_route_off[k], r1 = d->od_route_off[k + 1];
        if (r1 == r0) continue;  // "chuck out the OD pairs without paths": no init slot is consumed
        Lcg rnd{42}, srv{(long long)d->randseed};
        const double t_init = initvalue;
        initvalue += 0.00001;
        ++n_init;
        const double rate = d->od_rate[k];
        bool active = rate >= 1;  // ODaction::ODaction (B/od.cpp:30-46)
        double mu = active ? 3600.0 / rate : 0.0;
        double t;
        if (rate > 0) {  // ODpair::execute (B/od.cpp:355-368): first booking at urandom(1.2, 3600/rate)
            const double temp = 3600 / rate;
            t = t_init + (1.2 + (temp - 1.2) * rnd.urandom());
        } else {  // ... executes the (inactive) ODaction on the spot, which books a poll
            t = t_init + d->update_interval_routes + (1.0 + (100.0 - 1.0) * rnd.urandom());
        }
        int si = sl_off[(size_t)k];
        const int s1 = sl_off[(size_t)k + 1];
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
        double prev = t_init;  // the booking event: ODpair::execute, then the pair's previous event
        while (std::isfinite(t)) {  // (a rate of 0 books the next event at +inf: the pair is dead for good)
            for (; si < s1 && d->slice_time[sl_idx[(size_t)si]] <= t; ++si) {  // MatrixActions up to this instant come first
                const double r = d->slice_rate[sl_idx[(size_t)si]];
                mu = 3600 / r;  // ODaction::set_rate (B/od.h:63): +inf for a rate of 0
                active = true;
            }
            if (!active) {  // B/od.cpp:102-107: no vehicle; poll again after update_interval_routes + urandom(1, 100)
                if (t >= d->runtime) {
                    phantom(t, prev);
                    break;
                }
                prev = t;
                t = t + d->update_interval_routes + (1.0 + (100.0 - 1.0) * rnd.urandom());
                continue;
            }
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
    if (d->phantom_final) {
        d->phantom_final[0] = have_ph ? ph_t : 0.0;
        d->phantom_final[1] = have_ph ? ph_tp : 0.0;
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

// Vtypes::random_vtype per trip (B/vtypes.h:61-71, B/od.cpp:79): one network-wide Random stream, one draw per trip in global
// event order (include/mezzo_b200.h).
extern "C" int mz_vtypes_draw(int64_t randseed, int32_t n_types, const double *probability, int64_t n_trips, int32_t *type_index)
{
    MZ_REQUIRE(n_types >= 1 && probability && (type_index || n_trips == 0) && n_trips >= 0, MZ_EINVAL, "bad argument");
    MZ_REQUIRE(randseed != 0, MZ_EINVAL, "randseed must be != 0 (0 means clock seeding in the reference)");
    Lcg rnd{(long long)randseed};
    for (int64_t i = 0; i < n_trips; ++i) {
        const double r = rnd.urandom();
        double sum = 0.0;
        int32_t pick = n_types - 1;
        for (int32_t k = 0; k < n_types; ++k) {
            sum += probability[k];
            if (sum >= r) {
                pick = k;
                break;
            }
        }
        type_index[i] = pick;
    }
    return MZ_OK;
}

