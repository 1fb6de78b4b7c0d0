// Drop-in for hot path 1 inside mezzo_lib: the event loop of Network::step
//
//     while ((time>-1.0) && (time<runtime)) time=eventlist->next();          (B/network.cpp:7804-7807)
//
// replaced by one call that flattens the object graph Network::executemaster + Network::init built (links, sdfuncs,
// servers, turnings, origins, destinations, OD pairs with their routes, historical link times) into an MzNet, runs the
// horizon on the GPU through the C ABI (mz_trips_generate, mz_sim_create, mz_sim_run) and writes the results back
// into the objects Network::writeall reads: one ODpair::report row per arrived vehicle (B/od.cpp:379-384,
// B/vehicle.cpp:79-91) and Link::avgtimes / tmp_avg (B/link.cpp:118-141, 536-556). B/ = the reference's mezzo_lib/src.
//
// Usage in mezzo_lib (C++11, compile with -I<mezzo_b200>/include and link libmezzo_b200.so):
//     #include "network.h"
//     #include "NetworkStep.h"
//     double Network::step(double) { return mzhost::network_step(this, randseed); }
// The class members it reads are private/protected in the reference: either add `friend` declarations or, as the
// tested harness does (oracle/Makefile: _ref/mezzo_dropin_run), compile this translation unit with -fno-access-control.
//
// What the engine does not model is REFUSED (exception with the reason), never ignored: incidents, bus lines, boundary nodes / virtual links, several vehicle classes, give-ways, server types 3/4.
#pragma once
#include <algorithm>
#include <list>
#include <map>
#include <numeric>
#include <stdexcept>
#include <cmath>
#include <cstdlib>
#include <string>
#include <thread>
#include <vector>

#include "mezzo_b200.h"
#include "network.h"

namespace mzhost {

struct StepResult {
    mz_sim_stats stats;
    std::vector<MzExit> exits;  // per-vehicle link exits (filled when want_exit_log), link = dense link index
    std::vector<int> link_ids;  // dense link index -> link id
};

inline void mz_check(int rc, const char *what)
{
    if (rc != MZ_OK) throw std::runtime_error(std::string(what) + ": " + mz_last_error());
}

inline double network_step(Network *net, long seed, int device = 0, StepResult *res = nullptr, bool want_exit_log = false)
{
    for (Incident *inc : net->incidents)
        if (inc->info_start >= 0.0 && inc->info_stop >= 0.0)
            throw std::runtime_error("mezzo_b200: incidents with information broadcast (en-route switching) are not built");
    if (!net->buslines.empty()) throw std::runtime_error("mezzo_b200: bus lines (BusMezzo transit layer) are not built");
    if (!net->virtuallinkmap.empty()) throw std::runtime_error("mezzo_b200: virtual links / boundary nodes are not built");
    if (net->vehtypes.vtypes.empty()) throw std::runtime_error("mezzo_b200: vehiclemix.dat holds no vehicle class");
    if (seed == 0) throw std::runtime_error("mezzo_b200: an explicit random seed is required (seed 0 = clock-seeded in the reference)");

    // ---- dense indices in the orders the reference iterates (ascending ids)
    std::map<int, int> node_ix, link_ix, sd_ix, srv_ix;
    for (auto &kv : net->originmap) node_ix[kv.first] = 0;
    for (auto &kv : net->destinationmap) node_ix[kv.first] = 0;
    for (auto &kv : net->junctionmap) node_ix[kv.first] = 0;
    { int i = 0; for (auto &kv : node_ix) kv.second = i++; }
    { int i = 0; for (auto &kv : net->linkmap) link_ix[kv.first] = i++; }
    { int i = 0; for (auto &kv : net->sdfuncmap) sd_ix[kv.first] = i++; }
    { int i = 0; for (auto &kv : net->servermap) srv_ix[kv.first] = i++; }
    const int L = (int)link_ix.size(), P = net->nrperiods;

    std::vector<int32_t> link_length, link_lanes, link_sd, link_in, link_out;
    std::vector<double> link_hist((size_t)L * P);
    std::vector<Link *> links;
    for (auto &kv : net->linkmap) {
        Link *l = kv.second;
        const size_t li = links.size();
        links.push_back(l);
        link_length.push_back(l->length);
        link_lanes.push_back(l->nr_lanes);
        link_sd.push_back(sd_ix.at(l->sdfunc->get_id()));
        link_in.push_back(node_ix.at(l->in_node->get_id()));
        link_out.push_back(node_ix.at(l->out_node->get_id()));
        for (int p = 0; p < P; ++p) {
            auto it = l->histtimes->times.find(p);
            if (it == l->histtimes->times.end()) throw std::runtime_error("mezzo_b200: a link has no historical time for a period");
            link_hist[li * P + p] = it->second;
        }
    }
    std::vector<int32_t> sd_type;
    std::vector<double> sd_vfree, sd_vmin, sd_romax, sd_romin, sd_alpha, sd_beta;
    for (auto &kv : net->sdfuncmap) {
        Sdfunc *f = kv.second;
        DynamitSdfunc *dy = dynamic_cast<DynamitSdfunc *>(f);
        sd_type.push_back(dy ? 1 : 0);
        sd_vfree.push_back(f->vfree);
        sd_vmin.push_back(f->vmin);
        sd_romax.push_back(f->romax);
        sd_romin.push_back(f->romin);
        sd_alpha.push_back(dy ? dy->alpha : 0.0);
        sd_beta.push_back(dy ? dy->beta : 0.0);
    }
    std::vector<int32_t> srv_type;
    std::vector<double> srv_mu, srv_sd, srv_delay;
    for (auto &kv : net->servermap) {
        srv_type.push_back(kv.second->type);
        srv_mu.push_back(kv.second->mu);
        srv_sd.push_back(kv.second->sd);
        srv_delay.push_back(kv.second->delay);
    }
    std::vector<int32_t> turn_node, turn_server, turn_in, turn_out, turn_lookback;
    for (auto &kv : net->turningmap) {  // turningmap order = Network::init order
        Turning *t = kv.second;
        turn_node.push_back(node_ix.at(t->node->get_id()));
        turn_server.push_back(srv_ix.at(t->server->get_id()));
        turn_in.push_back(link_ix.at(t->inlink->get_id()));
        turn_out.push_back(link_ix.at(t->outlink->get_id()));
        turn_lookback.push_back(t->size);
    }
    std::vector<int32_t> dest_node, dest_server, origin_node;
    std::map<int, int> origin_ix;
    for (auto &kv : net->destinationmap) {
        dest_node.push_back(node_ix.at(kv.first));
        Server *s = kv.second->server;
        dest_server.push_back(s && srv_ix.count(s->get_id()) ? srv_ix.at(s->get_id()) : -1);
    }
    for (auto &kv : net->originmap) {
        origin_ix[kv.first] = (int)origin_node.size();
        origin_node.push_back(node_ix.at(kv.first));
    }

    // ---- OD pairs (already sorted by od_less_than, pairs without routes already dropped by Network::init) and routes
    std::vector<ODpair *> &ods = net->odpairs;
    std::vector<double> od_rate;
    std::vector<int32_t> od_route_off(1, 0), od_routes, route_off(1, 0), route_links, route_id, route_meters;
    std::vector<double> route_sumcost0, route_last0;  // Route::cost cache state (set by delete_spurious_routes, if it ran)
    for (ODpair *od : ods) {
        od_rate.push_back(od->rate);
        for (Route *r : od->routes) {
            od_routes.push_back((int32_t)route_id.size());
            route_id.push_back(r->get_id());
            route_sumcost0.push_back(r->sumcost);
            route_last0.push_back(r->last_calc_time);
            int meters = 0;
            for (Link *l : r->links) {
                route_links.push_back(link_ix.at(l->get_id()));
                meters += l->length;
            }
            route_meters.push_back(meters);
            route_off.push_back((int32_t)route_links.size());
        }
        od_route_off.push_back((int32_t)od_routes.size());
    }

    // ---- trip table: the ODaction chain as a host pre-pass
    MzDemand dem{};
    dem.n_odpairs = (int32_t)ods.size();
    dem.n_routes = (int32_t)route_id.size();
    dem.n_periods = P;
    dem.n_turnings = (int32_t)turn_node.size();
    dem.od_rate = od_rate.data();
    dem.od_route_off = od_route_off.data();
    dem.od_routes = od_routes.data();
    dem.route_off = route_off.data();
    dem.route_links = route_links.data();
    dem.link_hist = link_hist.data();
    dem.periodlength = net->periodlength;
    dem.runtime = net->runtime;
    dem.update_interval_routes = theParameters->update_interval_routes;
    dem.kirchoff_alpha = theParameters->kirchoff_alpha;
    dem.od_servers_deterministic = theParameters->od_servers_deterministic ? 1 : 0;
    dem.n_signal_controls = (int32_t)net->signalcontrols.size();
    dem.route_sumcost0 = route_sumcost0.data();
    dem.route_last0 = route_last0.data();
    dem.randseed = seed;
    // demand slices: the MatrixActions booked by Network::readrates, in execution order (load time, then file order); each
    // rate goes to the FIRST OD pair with these ids (find_if in MatrixAction::execute, B/network.cpp:8213-8229)
    std::vector<int32_t> slice_od;
    std::vector<double> slice_time, slice_rate;
    {
        std::vector<std::pair<double, ODSlice *>> sl(net->odmatrix.slices.begin(), net->odmatrix.slices.end());
        std::stable_sort(sl.begin(), sl.end(), [](const std::pair<double, ODSlice *> &a, const std::pair<double, ODSlice *> &b) { return a.first < b.first; });
        for (auto &ts : sl)
            for (ODRate &r : ts.second->rates) {
                int32_t k = -1;
                for (size_t i = 0; i < ods.size() && k < 0; ++i)
                    if (ods[i]->odids() == r.odid) k = (int32_t)i;
                if (k < 0) throw std::runtime_error("mezzo_b200: a demand slice names an OD pair that does not exist");
                slice_od.push_back(k);
                slice_time.push_back(ts.first);
                slice_rate.push_back((double)r.rate);
            }
    }
    dem.n_slice_rates = (int32_t)slice_od.size();
    dem.slice_od = slice_od.data();
    dem.slice_time = slice_time.data();
    dem.slice_rate = slice_rate.data();
    double phantom_final[2] = {0.0, 0.0};
    dem.phantom_final = phantom_final;
    int64_t n_trips = 0;
    int32_t n_init = 0;
    mz_check(mz_trips_generate(&dem, 0, nullptr, nullptr, nullptr, nullptr, &n_trips, &n_init), "mz_trips_generate");
    std::vector<double> trip_time((size_t)n_trips + 1), trip_prev((size_t)n_trips + 1), trip_length((size_t)n_trips + 1);
    std::vector<int32_t> trip_od((size_t)n_trips + 1), trip_route((size_t)n_trips + 1), trip_origin((size_t)n_trips + 1);
    mz_check(mz_trips_generate(&dem, n_trips, trip_time.data(), trip_prev.data(), trip_od.data(), trip_route.data(), &n_trips,
                               &n_init), "mz_trips_generate");
    // the vehicle class of every trip: ODaction::execute draws it from the network's one Vtypes stream (B/od.cpp:79); the
    // classes are taken in the order Vtypes::initialize left them in, with the probabilities it normalised
    std::vector<int32_t> trip_vtype((size_t)n_trips + 1, 0);
    {
        std::vector<double> prob;
        for (Vtype *vt : net->vehtypes.vtypes) prob.push_back(vt->probability);
        if (prob.size() > 1)
            mz_check(mz_vtypes_draw(seed, (int32_t)prob.size(), prob.data(), n_trips, trip_vtype.data()), "mz_vtypes_draw");
    }
    for (int64_t v = 0; v < n_trips; ++v) {
        trip_origin[v] = origin_ix.at(ods[trip_od[v]]->get_origin()->get_id());
        trip_length[v] = net->vehtypes.vtypes[(size_t)trip_vtype[v]]->length;
    }

    MzNet mn{};
    mn.n_nodes = (int32_t)node_ix.size();
    mn.n_links = L;
    mn.n_sdfuncs = (int32_t)sd_type.size();
    mn.n_servers = (int32_t)srv_type.size();
    mn.n_turnings = (int32_t)turn_node.size();
    mn.n_dests = (int32_t)dest_node.size();
    mn.n_origins = (int32_t)origin_node.size();
    mn.n_routes = (int32_t)route_id.size();
    mn.n_trips = (int32_t)n_trips;
    mn.n_periods = P;
    mn.n_odpairs = n_init;
    mn.standard_veh_length = theParameters->standard_veh_length;
    mn.server_type = theParameters->server_type;
    mn.use_giveway = theParameters->use_giveway ? 1 : 0;
    mn.randseed = seed;
    mn.runtime = net->runtime;
    mn.periodlength = net->periodlength;
    mn.link_length = link_length.data();
    mn.link_lanes = link_lanes.data();
    mn.link_sdfunc = link_sd.data();
    mn.link_in_node = link_in.data();
    mn.link_out_node = link_out.data();
    mn.link_hist = link_hist.data();
    mn.sd_type = sd_type.data();
    mn.sd_vfree = sd_vfree.data();
    mn.sd_vmin = sd_vmin.data();
    mn.sd_romax = sd_romax.data();
    mn.sd_romin = sd_romin.data();
    mn.sd_alpha = sd_alpha.data();
    mn.sd_beta = sd_beta.data();
    mn.srv_type = srv_type.data();
    mn.srv_mu = srv_mu.data();
    mn.srv_sd = srv_sd.data();
    mn.srv_delay = srv_delay.data();
    mn.turn_node = turn_node.data();
    mn.turn_server = turn_server.data();
    mn.turn_in = turn_in.data();
    mn.turn_out = turn_out.data();
    mn.turn_lookback = turn_lookback.data();
    mn.dest_node = dest_node.data();
    mn.dest_server = dest_server.data();
    mn.origin_node = origin_node.data();
    mn.route_off = route_off.data();
    mn.route_links = route_links.data();
    mn.trip_time = trip_time.data();
    mn.trip_origin = trip_origin.data();
    mn.trip_route = trip_route.data();
    mn.trip_length = trip_length.data();
    mn.trip_od = trip_od.data();
    mn.trip_prev_time = trip_prev.data();
    mn.phantom_final_t = phantom_final[0];
    mn.phantom_final_tp = phantom_final[1];
    // incidents without information broadcast, in the order their start events were booked (Network::readincident puts
    // every new Incident at the FRONT of the vector, B/network.cpp:6375)
    std::vector<int32_t> inc_link, inc_sdfunc;
    std::vector<double> inc_start, inc_stop;
    for (auto it = net->incidents.rbegin(); it != net->incidents.rend(); ++it) {
        inc_link.push_back(link_ix.at((*it)->lid));
        inc_sdfunc.push_back(sd_ix.at((*it)->sid));
        inc_start.push_back((*it)->start);
        inc_stop.push_back((*it)->stop);
    }
    mn.n_incidents = (int32_t)inc_link.size();
    mn.inc_link = inc_link.data();
    mn.inc_sdfunc = inc_sdfunc.data();
    mn.inc_start = inc_start.data();
    mn.inc_stop = inc_stop.data();

    // server-rate changes (ChangeRateAction objects of readserverrates, in file order)
    std::vector<int32_t> chg_server;
    std::vector<double> chg_time, chg_mu, chg_sd;
    for (ChangeRateAction *c : net->changerateactions) {
        chg_server.push_back(srv_ix.at(c->server->get_id()));
        chg_time.push_back(c->time);
        chg_mu.push_back(c->mu);
        chg_sd.push_back(c->sd);
    }
    mn.n_rate_changes = (int32_t)chg_server.size();
    mn.chg_server = chg_server.data();
    mn.chg_time = chg_time.data();
    mn.chg_mu = chg_mu.data();
    mn.chg_sd = chg_sd.data();

    // traffic signals: SignalControl -> SignalPlan -> Stage -> turnings, in the vectors' (file) order
    std::map<int, int> turn_ix;
    { int i = 0; for (auto &kv : net->turningmap) turn_ix[kv.first] = i++; }
    std::vector<int32_t> sigc_plan_off(1, 0), sigp_stage_off(1, 0), sigs_turn_off(1, 0), sigs_turns;
    std::vector<double> sigp_start, sigp_stop, sigs_duration;
    for (SignalControl *sc : net->signalcontrols) {
        for (SignalPlan *sp : sc->signalplans) {
            sigp_start.push_back(sp->start);
            sigp_stop.push_back(sp->stop);
            for (Stage *st : sp->stages) {
                sigs_duration.push_back(st->duration);
                for (Turning *t : st->turnings) sigs_turns.push_back(turn_ix.at(t->id));
                sigs_turn_off.push_back((int32_t)sigs_turns.size());
            }
            sigp_stage_off.push_back((int32_t)sigs_duration.size());
        }
        sigc_plan_off.push_back((int32_t)sigp_start.size());
    }
    mn.n_sig_controls = (int32_t)net->signalcontrols.size();
    mn.n_sig_plans = (int32_t)sigp_start.size();
    mn.n_sig_stages = (int32_t)sigs_duration.size();
    mn.sigc_plan_off = sigc_plan_off.data();
    mn.sigp_start = sigp_start.data();
    mn.sigp_stop = sigp_stop.data();
    mn.sigp_stage_off = sigp_stage_off.data();
    mn.sigs_duration = sigs_duration.data();
    mn.sigs_turn_off = sigs_turn_off.data();
    mn.sigs_turns = sigs_turns.data();

    // ---- the horizon on the GPU(s). MEZZO_B200_GPUS=N (N <= 8) partitions the network over the devices device .. device+N-1:
    // this one process creates an engine per GPU, maps their arenas into each other (direct peer access over NVLink) and
    // runs them from N host threads — the kernels synchronise pairwise through the arenas, there is no collective and no
    // second process. Afterwards the single event beyond the horizon goes to the rank that owns it and the
    // per-rank outputs are merged: a vehicle is reported by the rank of its destination node, a link's period averages by
    // the rank of its downstream node, a link exit by the rank of the exiting node.
    const char *gpus_env = std::getenv("MEZZO_B200_GPUS");
    const int n_gpus = gpus_env ? std::max(1, std::atoi(gpus_env)) : 1;
    std::vector<mz_sim *> sims((size_t)n_gpus, nullptr);
    struct Guard {
        std::vector<mz_sim *> &v;
        ~Guard() { for (mz_sim *s : v) if (s) mz_sim_destroy(s); }
    } guard{sims};
    mz_sim_stats st{};
    std::vector<double> arrival((size_t)n_trips + 1), avg((size_t)L * P + 1);
    if (n_gpus == 1) {
        mz_check(mz_sim_create(device, &mn, &sims[0]), "mz_sim_create");
        mz_check(mz_sim_run(sims[0]), "mz_sim_run");
        mz_check(mz_sim_last_stats(sims[0], &st), "mz_sim_last_stats");
        mz_check(mz_sim_get_arrivals(sims[0], arrival.data()), "mz_sim_get_arrivals");
        mz_check(mz_sim_get_linktimes_final(sims[0], avg.data()), "mz_sim_get_linktimes_final");
    } else {
        std::vector<int32_t> node_rank((size_t)mn.n_nodes);
        mz_check(mz_partition_nodes(&mn, n_gpus, node_rank.data()), "mz_partition_nodes");
        for (int r = 0; r < n_gpus; ++r)
            mz_check(mz_sim_create_part(device + r, &mn, n_gpus, r, node_rank.data(), &sims[(size_t)r]), "mz_sim_create_part");
        for (int r = 0; r < n_gpus; ++r)
            for (int q = 0; q < n_gpus; ++q)
                if (q != r) mz_check(mz_sim_attach_peer_local(sims[(size_t)r], q, sims[(size_t)q]), "mz_sim_attach_peer_local");
        std::vector<std::string> errs((size_t)n_gpus);
        {
            std::vector<std::thread> th;
            for (int r = 0; r < n_gpus; ++r)
                th.emplace_back([&, r] {
                    if (mz_sim_run(sims[(size_t)r]) != MZ_OK) errs[(size_t)r] = std::string("mz_sim_run, rank ") + std::to_string(r) + ": " + mz_last_error();
                });
            for (std::thread &t : th) t.join();
        }
        for (const std::string &e : errs)
            if (!e.empty()) throw std::runtime_error(e);
        // "while (time < runtime) time = next()": the first event at / beyond the horizon still executes — the smallest
        // pending key of all ranks, unless a MatrixAction or the poll of an inactive ODaction precedes it
        double bt = 0.0, btp = 0.0, end_time = 0.0;
        int32_t bsub = 0, bnode = -1;
        for (int r = 0; r < n_gpus; ++r) {
            double t, tp;
            int32_t sub, node;
            mz_check(mz_sim_min_key(sims[(size_t)r], &t, &tp, &sub, &node), "mz_sim_min_key");
            if (node < 0 || !std::isfinite(t)) continue;
            if (bnode < 0 || t < bt || (t == bt && (tp < btp || (tp == btp && (sub < bsub || (sub == bsub && node < bnode)))))) {
                bt = t; btp = tp; bsub = sub; bnode = node;
            }
        }
        if (bnode >= 0) end_time = bt;
        if (mn.phantom_final_t > 0.0 && (bnode < 0 || mn.phantom_final_t < bt || (mn.phantom_final_t == bt && mn.phantom_final_tp < btp))) {
            bnode = -1;
            end_time = mn.phantom_final_t;
        }
        std::vector<double> a((size_t)n_trips + 1), lt((size_t)L * P + 1);
        std::fill(arrival.begin(), arrival.end(), -1.0);
        std::fill(avg.begin(), avg.end(), 0.0);
        for (int r = 0; r < n_gpus; ++r) {
            mz_sim *s = sims[(size_t)r];
            mz_check(mz_sim_final_event(s, bnode, end_time), "mz_sim_final_event");  // (only the owner of bnode executes it)
            mz_sim_stats sr;
            mz_check(mz_sim_last_stats(s, &sr), "mz_sim_last_stats");
            st.n_traversals += sr.n_traversals; st.n_exits += sr.n_exits; st.n_events += sr.n_events; st.n_arrived += sr.n_arrived;
            st.n_supersteps += sr.n_supersteps; st.n_launches += sr.n_launches;
            st.kernel_ms = std::max(st.kernel_ms, sr.kernel_ms);
            st.total_ms = std::max(st.total_ms, sr.total_ms);
            mz_check(mz_sim_get_arrivals(s, a.data()), "mz_sim_get_arrivals");
            for (int64_t v = 0; v < n_trips; ++v) arrival[(size_t)v] = std::max(arrival[(size_t)v], a[(size_t)v]);
            mz_check(mz_sim_get_linktimes_final(s, lt.data()), "mz_sim_get_linktimes_final");  // zeros for links of other ranks
            for (size_t i = 0; i < (size_t)L * P; ++i) avg[i] += lt[i];
        }
        st.end_time = end_time;
    }
    mz_sim *sim = sims[0];

    // ---- results back into the objects Network::writeall reads
    // OD grids: rows in arrival order per OD pair (Vehicle::report: vid, start, end, travel time, meters, route id, switched;
    // ODpair::report prepends origin and destination)
    std::vector<int64_t> done;
    for (int64_t v = 0; v < n_trips; ++v)
        if (arrival[v] >= 0) done.push_back(v);
    std::stable_sort(done.begin(), done.end(), [&](int64_t a, int64_t b) { return arrival[a] < arrival[b]; });
    for (int64_t v : done) {
        const int r = trip_route[v];
        std::list<double> row;
        row.push_back((double)(v + 1));
        row.push_back(trip_time[v]);
        row.push_back(arrival[v]);
        row.push_back(arrival[v] - trip_time[v]);
        row.push_back((double)route_meters[r]);
        row.push_back((double)route_id[r]);
        row.push_back(0.0);
        ods[trip_od[v]]->report(row);
    }
    // link times: after Link::end_of_simulation the period averages must read avg[l][p]; the open period is carried by tmp_avg
    for (int l = 0; l < L; ++l) {
        Link *lk = links[l];
        for (int p = 0; p < P; ++p) lk->avgtimes->times[p] = avg[(size_t)l * P + p];
        if (lk->curr_period >= 0 && lk->curr_period < P) lk->tmp_avg = avg[(size_t)l * P + lk->curr_period];
    }
    net->time = st.end_time;
    if (res) {
        res->stats = st;
        for (auto &kv : net->linkmap) res->link_ids.push_back(kv.first);
        if (want_exit_log) {
            (void)sim;
            for (mz_sim *s : sims) {  // every rank logs the exits its own nodes performed
                int64_t n = 0;
                mz_check(mz_sim_get_exit_log(s, nullptr, 0, &n), "mz_sim_get_exit_log");
                const size_t at = res->exits.size();
                res->exits.resize(at + (size_t)n + 1);
                mz_check(mz_sim_get_exit_log(s, res->exits.data() + at, n, &n), "mz_sim_get_exit_log");
                res->exits.resize(at + (size_t)n);
            }
            if (sims.size() > 1)
                std::stable_sort(res->exits.begin(), res->exits.end(), [](const MzExit &a, const MzExit &b) {
                    return a.vid != b.vid ? a.vid < b.vid : a.entry < b.entry;
                });
        }
    }
    return st.end_time;
}

}  // namespace mzhost
