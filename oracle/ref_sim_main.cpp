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
    size_t n_on_links = 0;
    for (auto &kv : net->get_links()) n_on_links += (size_t)kv.second->size();
    f = fopen((out + "info.txt").c_str(), "w");
    fprintf(f, "step_seconds=%.9f\nn_exits=%zu\nn_trips=%zu\nn_arrivals=%zu\ncurrenttime=%.17g\nn_on_links=%zu\ninit_seconds=%.9f\n", secs,
            g_exits.size(), g_trips.size(), n_arr, net->get_currenttime(), n_on_links, init_secs);
#ifdef MZ_DROPIN
    if (net->graph) fprintf(f, "sssp_prefetch_batches=%ld\nsssp_prefetch_served=%ld\n", net->graph->prefetchBatches(), net->graph->prefetchServed());
#endif
    fclose(f);
    if (argc > 4 && std::string(argv[4]) == "save") nt->saveresults();
    fflush(stdout);
    _exit(0);  // skip the reference's destructors (double deletes on some paths)
}
