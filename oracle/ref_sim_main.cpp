// TEST INFRASTRUCTURE — harness around the UNMODIFIED reference mezzo_lib (compiled in place from /root/reference,
// see oracle/Makefile) = the oracle for hot path 1 (link update) and the CPU baseline of bench.py.
//
//   mezzo_ref_run <masterfile> <seed> <outprefix>
// runs NetworkThread exactly like B/main.cpp:42-50 (init → start → wait), timing only Network::step, and dumps
//   <outprefix>exits.bin     {int32 vid, int32 lid, double entry, double exit} per successful Link::exit_veh
//   <outprefix>trips.bin     {int32 vid, int32 route_id, int32 oid, int32 did, double time, double length} per ODaction
//   <outprefix>arrivals.bin  9 doubles per OD grid row (o,d,vid,start,end,tt,meters,route,switched), OD pairs in order
//   <outprefix>info.txt      key=value: step_seconds, n_exits, n_trips, n_arrivals, currenttime
#include <unistd.h>
#include <chrono>
#include <cstdio>
#include <fstream>
#include <vector>

#include "network.h"
#ifdef MZ_DROPIN_STEP
// mezzo_dropin_run: the SAME harness and the SAME reference library, but Network::step's event loop is replaced by the GPU
// engine through the product's C++ host glue (mezzo_b200/host/NetworkStep.h) — the drop-in proof for hot path 1.
#include "NetworkStep.h"
#endif

struct ExitRec { int vid, lid; double entry, exit; };
struct TripRec { int vid, route, oid, did; double time, length; };
static std::vector<ExitRec> g_exits;
static std::vector<TripRec> g_trips;

void mz_hook_exit(int vid, int lid, double entry_time, double exit_time)
{
    g_exits.push_back(ExitRec{vid, lid, entry_time, exit_time});
}

struct StopRec { int vid, trip, stop, link; double enter, exit; };
struct DispatchRec { int vid, trip; double time, length; };
static std::vector<StopRec> g_stops;
static std::vector<DispatchRec> g_dispatch;
void mz_hook_stop_visit(int vid, int trip_id, int stop_id, int link_id, double enter_time, double exit_time)
{
    g_stops.push_back(StopRec{vid, trip_id, stop_id, link_id, enter_time, exit_time});
}
void mz_hook_bus_dispatch(int vid, int trip_id, double dispatch_time, double vehicle_length)
{
    g_dispatch.push_back(DispatchRec{vid, trip_id, dispatch_time, vehicle_length});
}

void mz_hook_trip(Vehicle *veh, double time)
{
    odval od = veh->get_odids();
    g_trips.push_back(TripRec{veh->get_id(), veh->get_route()->get_id(), od.first, od.second, time,
                              (double)veh->get_length()});
}

int main(int argc, char **argv)
{
    if (argc < 4) {
        fprintf(stderr, "usage: %s masterfile seed outprefix\n", argv[0]);
        return 2;
    }
    long seed = atol(argv[2]);
    std::string out = argv[3];
    NetworkThread *nt = new NetworkThread(argv[1], 1, seed);
    auto ti0 = std::chrono::steady_clock::now();
    nt->init();  // readmaster + executemaster: with calc_paths=1 this is where Network::shortest_paths_all runs
    const double init_secs = std::chrono::duration<double>(std::chrono::steady_clock::now() - ti0).count();
    g_exits.reserve(1 << 20);
    auto t0 = std::chrono::steady_clock::now();
#ifdef MZ_DROPIN_STEP
    Network *net = nt->getNetwork();
    mzhost::StepResult res;
    try {
        mzhost::network_step(net, seed, 0, &res, true);
    } catch (const std::exception &e) {
        fprintf(stderr, "drop-in step failed: %s\n", e.what());
        _exit(3);
    }
    for (const MzExit &x : res.exits) g_exits.push_back(ExitRec{x.vid, res.link_ids[x.link], x.entry, x.exit});
#else
    nt->start(QThread::HighestPriority);
    nt->wait();
    Network *net = nt->getNetwork();
#endif
    double secs = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();

    FILE *f = fopen((out + "exits.bin").c_str(), "wb");
    fwrite(g_exits.data(), sizeof(ExitRec), g_exits.size(), f);
    fclose(f);
    f = fopen((out + "trips.bin").c_str(), "wb");
    fwrite(g_trips.data(), sizeof(TripRec), g_trips.size(), f);
    fclose(f);
    // OD grids (ODpair::report → Grid rows, B/od.cpp:379-384); private members via -fno-access-control
    size_t n_arr = 0;
    f = fopen((out + "arrivals.bin").c_str(), "wb");
    for (ODpair *od : net->odpairs) {
        for (auto &row : od->grid->grid) {
            std::vector<double> r(row.begin(), row.end());
            r.resize(9, 0.0);
            fwrite(r.data(), sizeof(double), 9, f);
            ++n_arr;
        }
    }
    fclose(f);
    // vehicles sitting on a real link when the loop ended: with the exits they make up the successful Link::enter_veh
    // calls on real links, i.e. the traversals of bench.py's metric
    // transit layer (BusMezzo fixtures): the stop visits and dispatches the hooks saw, and the static description of every
    // bus trip — vehicle, length, route links, stops (link, position) — for tests/golden/make_bus_golden.py
    f = fopen((out + "stops.bin").c_str(), "wb");
    if (!g_stops.empty()) fwrite(g_stops.data(), sizeof(StopRec), g_stops.size(), f);
    fclose(f);
    f = fopen((out + "busdispatch.bin").c_str(), "wb");
    if (!g_dispatch.empty()) fwrite(g_dispatch.data(), sizeof(DispatchRec), g_dispatch.size(), f);
    fclose(f);
    f = fopen((out + "bustrips.txt").c_str(), "w");
    for (Bustrip *bt : net->bustrips) {
        Bus *bus = bt->get_busv();
        Busline *line = bt->get_line();
        fprintf(f, "trip %d line %d vid %d length %.17g starttime %.17g origin %d dest %d route %d links", bt->get_id(), line->get_id(),
                bus ? bus->get_id() : -1, bus ? (double)bus->get_length() : 0.0, bt->get_starttime(), line->get_odpair()->get_origin()->get_id(),
                line->get_odpair()->get_destination()->get_id(), line->get_busroute()->get_id());
        for (Link *l : line->get_busroute()->get_links()) fprintf(f, " %d", l->get_id());
        fprintf(f, " stops");
        for (Visit_stop *vs : bt->stops) fprintf(f, " %d:%d:%.17g", vs->first->get_id(), vs->first->get_link_id(), vs->first->get_position());
        fprintf(f, "\n");
    }
    fclose(f);
    size_t n_on_links = 0;
    for (auto &kv : net->get_links()) n_on_links += (size_t)kv.second->size();
    f = fopen((out + "info.txt").c_str(), "w");
    fprintf(f, "step_seconds=%.9f\nn_exits=%zu\nn_trips=%zu\nn_arrivals=%zu\ncurrenttime=%.17g\nn_on_links=%zu\ninit_seconds=%.9f\n", secs,
            g_exits.size(), g_trips.size(), n_arr, net->get_currenttime(), n_on_links, init_secs);
    {   // the 0.00001 steps the transit layer adds to the init clock (B/network.cpp:7672-7694)
        int steps = (int)net->buslines.size();
        if (theParameters->demand_format == 3)
            for (ODstops *od : net->odstops_demand)
                if (od->get_arrivalrate() != 0.0) ++steps;
        fprintf(f, "n_transit_init_steps=%d\n", steps);
    }
#ifdef MZ_DROPIN
    if (net->graph) fprintf(f, "sssp_prefetch_batches=%ld\nsssp_prefetch_served=%ld\n", net->graph->prefetchBatches(), net->graph->prefetchServed());
#endif
    fclose(f);
    if (argc > 4 && std::string(argv[4]) == "save") nt->saveresults();
    fflush(stdout);
    _exit(0);  // skip the reference's destructors (double deletes on some paths)
}
