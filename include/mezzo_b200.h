/*
 * mezzo_b200.h — C ABI of the B200-native engine for Mezzo's two data-parallel hot paths.
 *
 * Citations use B/ = /root/reference/branches/busmezzo/mezzo_lib/src/.
 *
 * The reference has no FFI/plugin seam; the seam is the C++ class API of mezzo_lib.  The entry
 * points below are what a mezzo_lib maintainer binds at the two call sites that are replaced
 * (the binding code is shown in INTEGRATION.md and shipped in mezzo_b200/host/):
 *
 *   hot path 2 (route generation)   graph->labelCorrecting(root, entry, linkinfo)   B/network.cpp:6714
 *                                   graph->reachable(d) / get_path(d)               B/network.cpp:6726-6731
 *   hot path 1 (link update)        while (time<runtime) time=eventlist->next();    B/network.cpp:7804-7807
 *
 * Rules (all entry points):
 *   - extern "C", plain pointers and sizes; no C++/torch types cross the boundary.
 *   - return 0 (MZ_OK) on success, a negative MZ_E* code on error; never throws, never aborts.
 *     mz_last_error() returns a human-readable string for the calling thread's last error.
 *   - the caller owns every host buffer it passes; the engine owns all device memory behind the
 *     opaque handle.  Output pointers may be host (pageable or pinned) or device pointers: the
 *     engine inspects them (cudaPointerGetAttributes) and skips the staging copy for device memory.
 *   - one calling thread per handle (mezzo_lib itself is non-reentrant, B/main.cpp:40-41).
 *   - there is NO CPU fallback: without a CUDA device every compute entry point returns MZ_ENODEV.
 */
#ifndef MEZZO_B200_H
#define MEZZO_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* B/Graph.h:27 */
#define MZ_SP_UNSET (-2)
#define MZ_SP_START (-1)
/* B/network.cpp:6598 (Graph ctor 3rd arg) */
#define MZ_SP_INFINITY 9999999.0
/* B/linktimes.cpp:74-80,119 : "NEVER RETURN 0" */
#define MZ_MISSING_COST 0.1

enum {
    MZ_OK = 0,
    MZ_EINVAL = -1,   /* bad argument (null pointer, negative size, id out of range ...) */
    MZ_ENODEV = -2,   /* no CUDA device / driver; the engine never falls back to the CPU */
    MZ_ECUDA = -3,    /* a CUDA runtime call failed; see mz_last_error() */
    MZ_ENOMEM = -4,   /* host or device allocation failed */
    MZ_ESTATE = -5,   /* call order violated (e.g. sssp before linktimes_set) */
    MZ_ELIMIT = -6    /* a structural limit of the reference data model is exceeded */
};

typedef struct mz_graph mz_graph;   /* opaque: route-generation graph + link-time table on one GPU */
typedef struct mz_sim mz_sim;       /* opaque: link-update engine state on one GPU */

const char *mz_last_error(void);
int mz_version(void);
/* number of visible CUDA devices (0 if none / no driver); never fails */
int mz_device_count(void);

/* ------------------------------------------------------------------------------------------------
 * Hot path 2: route generation  (Graph<double,LinkTimeInfo>, B/Graph.h, B/Graph.cpp)
 * ------------------------------------------------------------------------------------------------ */

/*
 * Replaces: new Graph<double,LinkTimeInfo>(max node id+1, max link id+1, 9999999.0) followed by
 * graph->addLink(lid,in,out,cost) for each link, graph->set_downlink_indices() and
 * graph->set_turning_prohibitor(from,to)                      B/network.cpp:6598-6625, B/Graph.cpp:47-63,689-717
 *
 *   device         CUDA device ordinal the graph lives on.
 *   n_node_slots   = max node id + 1;  n_link_slots = max link id + 1  (ids are array indices).
 *   n_added        number of addLink calls; add_order[k] = link id of the k-th call (NULL ⇒ ascending
 *                  id over the slots whose up_node != MZ_SP_UNSET, which is what Network does by iterating linkmap).
 *                  The order fixes dnLinks_/upLinks_ order per node and hence dnIndex_.
 *   up_node/dn_node[n_link_slots]   end nodes per slot, MZ_SP_UNSET for slots never added.
 *   dn_legal[n_link_slots]          the char dnLegal_ bitmask per link AFTER prohibitors were XOR-ed in
 *                                   (NULL ⇒ 0xFF everywhere). Bit k refers to dnIndex k of the out-link; the reference
 *                                   truncates the test to 8 bits (B/Graph.h:127-129) so out-links with dnIndex>=8 are never relaxed.
 * Errors: MZ_ELIMIT if a node has more than 127 out-links (dnIndex_ is a char in the reference).
 */
int mz_graph_create(int device, int32_t n_node_slots, int32_t n_link_slots, int32_t n_added,
                    const int32_t *add_order, const int32_t *up_node, const int32_t *dn_node,
                    const int8_t *dn_legal, mz_graph **out);

/*
 * Replaces: LinkTimeInfo / LinkTime (B/linktimes.h:34-66) as consumed by LinkTimeInfo::cost (B/linktimes.cpp:108-121).
 *   T[n_link_slots * n_periods] row-major by link slot; NaN = "period not present" (cost 0.1, B/linktimes.cpp:74-80).
 *   cost(l,t) = T[l][int(t/periodlength)] if that index is inside [0,n_periods) and not NaN, else 0.1.
 *   T may be a host or device pointer. The table is replicated on the graph's device.
 */
int mz_linktimes_set(mz_graph *g, int32_t n_periods, double periodlength, const double *T);

/*
 * Replaces: n calls of graph->labelCorrecting(root[i], entry[i], linkinfo)          B/Graph.cpp:313-454
 * including calcCostToNodes (B/Graph.cpp:458-483).
 *   link_pred[n * n_link_slots]  (int32)  GraphLink::predecessor_ per slot per tree   (may be NULL)
 *   node_pred[n * n_node_slots]  (int32)  GraphNode::predecessor_                      (may be NULL)
 *   node_cost[n * n_node_slots]  (double) GraphNode::cost_ (9999999.0 = unreachable)   (may be NULL)
 * Bit-exact with the reference for every input (exact deque emulation on the device);
 * mode selects the kernel:  MZ_SSSP_EXACT always emulates the sequential deque;
 *                           MZ_SSSP_AUTO uses the batched frontier kernel and certifies each tree on the device
 *                           (FIFO table + unique argmin per link); uncertified trees are redone by the exact kernel.
 */
enum { MZ_SSSP_EXACT = 0, MZ_SSSP_AUTO = 1 };
int mz_sssp_batch(mz_graph *g, int32_t n, const int32_t *root, const double *entry, int mode,
                  int32_t *link_pred, int32_t *node_pred, double *node_cost);

/*
 * Replaces the per-destination part of Network::shortest_paths_all: graph->reachable(dest) and
 * Network::get_path → Graph::shortest_path_vector (B/network.cpp:6726-6742, 6633-6664, B/Graph.cpp:653-685),
 * including "prepend the root link if it is not the first link" (B/network.cpp:6741-6742).
 * Trees are computed on the device as in mz_sssp_batch and walked there; only the paths come back.
 *   dest_off[n+1], dest[dest_off[n]]   CSR of destination NODE ids per tree.
 *   path_off[dest_off[n]+1]            out: CSR offsets into path_links per (tree,dest) pair, in input order.
 *   path_links[path_cap]               out: link ids root→dest; an unreachable dest or a circular path yields length 0.
 *   n_path_links                       out: total links written (if > path_cap: MZ_ELIMIT, nothing written beyond cap).
 */
int mz_sssp_routes(mz_graph *g, int32_t n, const int32_t *root, const double *entry, int mode,
                   const int64_t *dest_off, const int32_t *dest, int64_t *path_off,
                   int32_t *path_links, int64_t path_cap, int64_t *n_path_links);

/* Counters of the last mz_sssp_batch / mz_sssp_routes call (metric 2):
 *   canonical_relax = Σ_trees Σ_{reached u} |legal out-links of dnNode(u) not pointing to the root node|
 *   evaluated_relax = relaxations the device actually evaluated
 *   kernel_ms       = device time of the SSSP kernels of that call (CUDA events on the engine stream)
 *   n_exact_redo    = trees the AUTO mode had to redo with the exact kernel */
typedef struct {
    int64_t canonical_relax;      /* -1 if counting was switched off (MZ_OPT_COUNT_CANONICAL) */
    int64_t evaluated_relax;
    int64_t reached_links;        /* Σ_trees |{links with a predecessor}| (-1 if counting is off) */
    double kernel_ms;             /* all kernels of the call (init + main + node preds + counters + route walk) */
    double main_kernel_ms;        /* the tree-building kernel(s) only */
    double total_ms;              /* host wall time of the call, copies included */
    int32_t n_exact_redo;
    int32_t n_launches;           /* kernels launched by the call */
    int32_t n_chunks;             /* tree chunks the batch was split into (device work space is bounded) */
    int32_t main_kernel_launches;
    int64_t fast_steps;           /* near-far steps of the batched kernel (0 in exact mode) */
} mz_sssp_stats;
int mz_sssp_last_stats(const mz_graph *g, mz_sssp_stats *out);

/* Tuning knobs (never change results):
 *   MZ_OPT_CHUNK_TREES      max trees resident per chunk (0 = size from free device memory)
 *   MZ_OPT_COUNT_CANONICAL  0 switches the canonical-relaxation counter kernel off
 *   MZ_OPT_DELTA_MILLI      bucket width of the batched (MZ_SSSP_AUTO) kernel in thousandths of the mean link cost
 *   MZ_OPT_MAX_STEPS        step cap of the batched kernel; trees unfinished at the cap are redone by the exact kernel */
enum { MZ_OPT_CHUNK_TREES = 1, MZ_OPT_COUNT_CANONICAL = 2, MZ_OPT_DELTA_MILLI = 3, MZ_OPT_MAX_STEPS = 4 };
int mz_graph_set_option(mz_graph *g, int option, int64_t value);

/* Run the engine's kernels on a caller-owned CUDA stream (cudaStream_t passed as void*; NULL restores the
 * engine's own stream). Lets the caller bracket calls with its own CUDA events on that stream. */
int mz_graph_set_stream(mz_graph *g, void *cuda_stream);

int mz_graph_destroy(mz_graph *g);


/* ------------------------------------------------------------------------------------------------
 * Hot path 1: link update  (Network::step event loop over Turning/Link/Q/Sdfunc/Server)
 * ------------------------------------------------------------------------------------------------ */

/*
 * MzNet — the flattened network a mezzo_lib host hands over after Network::executemaster (B/network.cpp:7296-7390).
 * Everything is structure-of-arrays with DENSE 0-based indices (the host keeps the id maps); no per-vehicle heap
 * objects exist on the engine side. Orders are part of the contract because they fix the reference's event order:
 *   turnings      in turningmap order (ascending turning id)        B/network.cpp:7661-7665
 *   destinations  in destinationmap order (ascending node id)       B/network.cpp:7727-7731
 *   trips         in global ODaction execution order; vid = index+1 B/od.cpp:69-108 (host pre-pass)
 */
typedef struct {
    int32_t n_nodes, n_links, n_sdfuncs, n_servers, n_turnings, n_dests, n_origins, n_routes, n_trips, n_periods;
    int32_t n_odpairs;            /* OD pairs initialised by Network::init (those with >= 1 route): offsets the destinations' init times;
                                     trip_od values are in [0, n_odpairs) */
    int32_t standard_veh_length;  /* Parameters::standard_veh_length (int), maxcap = int(length*lanes/standard_veh_length)  B/link.cpp:11 */
    int32_t server_type;          /* Parameters::server_type: global override inside Server::next  B/server.cpp:32-50: 3 log-normal, 4 log-logistic headways */
    int32_t use_giveway;          /* must be 0 (B/turning.cpp:147-150) */
    int64_t randseed;             /* every Random is seeded with it (B/server.cpp:9-13); must be != 0 */
    double runtime;               /* stoptime; events run while time < runtime, plus the first event beyond (B/network.cpp:7804-7806) */
    double periodlength;          /* link-time period length of histtimes/avgtimes (B/network.cpp:6239-6241) */
    /* links */
    const int32_t *link_length, *link_lanes, *link_sdfunc, *link_in_node, *link_out_node; /* [n_links] */
    const double *link_hist;      /* [n_links*n_periods] histtimes->times (fallback of the period averages, B/link.cpp:541-545) */
    /* speed-density functions: type 0 Sdfunc, 1|2 DynamitSdfunc   B/sdfunc.cpp:12-53 */
    const int32_t *sd_type;
    const double *sd_vfree, *sd_vmin, *sd_romax, *sd_romin, *sd_alpha, *sd_beta;
    /* servers: type 0 ConstServer, 1 Server N(mu,sd)>=0.1, 2 DetServer   B/server.h:43-99 */
    const int32_t *srv_type;
    const double *srv_mu, *srv_sd, *srv_delay;
    /* turnings */
    const int32_t *turn_node, *turn_server, *turn_in, *turn_out, *turn_lookback; /* [n_turnings] */
    /* destinations; dest_server -1 = the node's own default server N(0.1,0.5), delay 0.001  B/node.cpp:300-304 */
    const int32_t *dest_node, *dest_server; /* [n_dests] */
    const int32_t *origin_node;             /* [n_origins] */
    /* routes */
    const int32_t *route_off;   /* [n_routes+1] */
    const int32_t *route_links; /* link indices */
    /* trips */
    const double *trip_time;    /* ODaction execution time = vehicle start time */
    const int32_t *trip_origin; /* index into origin_node */
    const int32_t *trip_route;
    const double *trip_length;  /* vehicle length (m) */
    const int32_t *trip_od;     /* OD pair (ODaction) that generates the trip: its events are FIFO-ordered like any action's */
    const double *trip_prev_time; /* time of the event that booked this trip: the same OD pair's previous trip, or
                                     ODpair::execute's init time for its first trip (B/od.cpp:355-368) */
    /* The trips of one trip_od value form ONE event chain — every trip is booked when the chain's previous trip executes —
     * and the trip table as a whole is in global event order: by trip_time, equal times by trip_prev_time, then by chain.
     * (A vehicle source that is an action of its own in the reference — a Busline dispatching its Bustrips,
     * B/busline.cpp:83-112 — gets a trip_od value of its own, [0, n_odpairs), not the one of an OD pair that happens to share
     * its origin and destination: trips of equal time filed in an order their chains do not execute in would overtake each
     * other in the origin's virtual queue.) */
    /* server-rate changes (serverrates.dat → ChangeRateAction, B/network.cpp:5608-5665, B/server.cpp:146-168): from
     * chg_time on the server's mu and/or sd are replaced (a value <= 0 leaves that parameter alone). A change is booked
     * before every action, so it precedes the events of its own time instant; equal times apply in array order. */
    int32_t n_rate_changes;
    const int32_t *chg_server;    /* [n_rate_changes] server index */
    const double *chg_time, *chg_mu, *chg_sd;
    /* traffic signals (signal.dat → SignalControl / SignalPlan / Stage, B/network.cpp:5144-5295, B/trafficsignal.cpp):
     * controls own plans, plans own stages, a stage lists the turnings it turns green (they start red). All turnings of
     * one control must belong to one node, and a turning to at most one control (MZ_ELIMIT otherwise). */
    int32_t n_sig_controls, n_sig_plans, n_sig_stages;
    const int32_t *sigc_plan_off;  /* [n_sig_controls+1] plans of a control, file order */
    const double *sigp_start, *sigp_stop; /* [n_sig_plans] */
    const int32_t *sigp_stage_off; /* [n_sig_plans+1] stages of a plan, file order */
    const double *sigs_duration;   /* [n_sig_stages] */
    const int32_t *sigs_turn_off;  /* [n_sig_stages+1] */
    const int32_t *sigs_turns;     /* turning indices, file order */
    /* The earliest event at/after `runtime` that is none of the network's actions — a MatrixAction (demand slice) or the
     * poll of an inactive ODaction (B/od.cpp:102-107), as mz_trips_generate reports it: (time, time of the event that booked
     * it; -inf for a MatrixAction, which is booked while the demand file is read). If it precedes every pending action it is
     * the one event Network::step still executes beyond the horizon (B/network.cpp:7804-7806) and no action runs there.
     * 0 = there is none. */
    double phantom_final_t, phantom_final_tp;
    /* Incidents WITHOUT information broadcast (incident file, Network::readincident B/network.cpp:6337-6380 with info_start < 0
     * or info_stop < 0; Incident::execute B/network.cpp:8092-8112): at inc_start the link's speed-density function is replaced
     * by inc_sdfunc (Link::set_incident, B/link.cpp:854-860 — its `blocked_until = -2.0` assigns the parameter, not the member),
     * at inc_stop the saved function comes back and the link is unblocked, blocked_until = -1.0 (Link::unset_incident,
     * B/link.cpp:862-868). The start events are booked while the incident file is read, which executemaster does right after
     * Network::init (B/network.cpp:7380-7383): they follow every booking of init and precede every booking of the run.
     * inc_sdfunc indexes the sd_* arrays (the incident file's own sdfuncs are part of the network's sdfunc map). */
    int32_t n_incidents;
    const int32_t *inc_link, *inc_sdfunc; /* [n_incidents] */
    const double *inc_start, *inc_stop;   /* 0 < start < stop */
    /* Transit overlay hooks (all NULL without buses). A trip with stops is a type-4 vehicle (Bus): when
     * Link::enter_veh puts it on the link of its next stop it is NOT queued — the engine reports a stop visit instead
     * (B/link.cpp:489-516) — and it comes back into that link's queue through Q::enter_veh when the host's transit model
     * (Busline / Bustrip / Busstop, passengers, dwell times: all of it stays on the host) lets it leave the stop
     * (Busstop::execute, B/busline.cpp:1196-1263). See mz_sim_run_until / mz_sim_stop_arrivals / mz_sim_stop_departures. */
    const int32_t *trip_stop_off; /* [n_trips+1] CSR: the stops of each trip in visit order (Bustrip::stops) */
    const int32_t *stop_link;     /* link of each stop (Busstop::link_id); the links must follow the trip's route order */
    const double *stop_pos;       /* Busstop::position, metres from the start of the link, 0 <= pos <= length */
    /* Network::init advances its init clock by 0.00001 for every Busline and for every ODstops pair with a non-zero arrival
     * rate (demand_format 3) between the signal controls and the OD pairs (B/network.cpp:7672-7694): the first polls of the
     * destinations' Dactions (and the OD pairs' init times) shift with it. The count of those steps. */
    int32_t n_transit_init_steps;
} MzNet;

/* One bus on its way to a stop: Bustrip::book_stop_visit(time_to_stop) as Link::enter_veh (first stop on a link,
 * B/link.cpp:503-511) or Busstop::execute (next stop on the same link, B/busline.cpp:1229-1236) computes it. */
typedef struct {
    int32_t veh;      /* trip index */
    int32_t stop;     /* index into the trip's own stop list */
    int32_t link;     /* link index */
    int32_t pad;
    double t_enter;   /* Vehicle::entry_time on that link */
    double t_arrive;  /* time_to_stop: when Busstop::execute sees the bus enter the stop */
} mz_stop_visit;

/* one record per successful Link::exit_veh (B/link.cpp:528-655): the per-vehicle link exit times of the parity contract */
typedef struct {
    int32_t vid;    /* trip index + 1 */
    int32_t link;   /* link index */
    double entry;   /* Vehicle::entry_time on that link */
    double exit;    /* time of the exit event */
} MzExit;

typedef struct {
    int64_t n_traversals;  /* successful Link::enter_veh on real links = metric 1 */
    int64_t n_exits;       /* successful Link::exit_veh (turnings + sinks) */
    int64_t n_events;      /* executed TurnAction / Daction / ODaction events */
    int64_t n_arrived;     /* vehicles that reached their destination */
    int64_t n_supersteps;  /* node visits that executed at least one event (the engine has no global steps) */
    int64_t n_launches;    /* kernels launched */
    double kernel_ms;      /* device time of the run (CUDA events on the engine stream) */
    double total_ms;       /* host wall time of mz_sim_run */
    double end_time;       /* Network::time after the loop (time of the last executed event) */
    int64_t n_stop_visits; /* stop visits reported so far (transit overlay hooks) */
    int64_t n_bus_refused; /* buses whose Q::enter_veh after a stop found the queue full (B/q.cpp:52-56 returns false:
                              the reference loses the vehicle, so does the engine) */
} mz_sim_stats;

/*
 * Host pre-pass for the trip table of MzNet. Replaces the ODaction event chain of the reference —
 * ODpair::execute's first booking (B/od.cpp:355-368), ODaction::execute / ODServer::next (B/od.cpp:69-108,
 * B/server.cpp:82-91) and ODpair::select_route over Route::cost (B/od.cpp:247-300, B/route.cpp:173-197) — which does not
 * depend on the traffic state. Pure host code: no CUDA device is needed or used.
 *   OD pairs are given in od_less_than order (origin id, destination id); pairs without routes consume no init slot.
 *   Outputs, in global event order (time; equal times by the booking event's time, then OD order): the trip's ODaction
 *   time, the time of the event that booked it, its OD pair (index into the input) and its route (index into routes).
 *   Call with cap = 0 and null buffers to obtain the trip count in *n_out first.
 */
typedef struct MzDemand {
    int32_t n_odpairs, n_routes, n_periods, n_turnings;  /* n_turnings: Turning::init calls before the OD pairs (init times) */
    const double *od_rate;        /* [n_odpairs] vehicles per hour (already scaled) */
    const int32_t *od_route_off;  /* [n_odpairs+1] CSR into od_routes */
    const int32_t *od_routes;     /* route indices of every pair, in the order ODpair::routes holds them */
    const int32_t *route_off;     /* [n_routes+1] CSR into route_links */
    const int32_t *route_links;   /* dense link indices */
    const double *link_hist;      /* [n_links * n_periods] historical link times (LinkTime::cost) */
    double periodlength, runtime;
    double update_interval_routes, kirchoff_alpha;
    int32_t od_servers_deterministic;
    int32_t n_signal_controls;    /* SignalControl::execute calls of Network::init between the turnings and the OD pairs
                                     (each advances the init time by 0.000001, B/network.cpp:7666-7670) */
    int64_t randseed;
    /* Route::sumcost / last_calc_time at the start of the run, per route; NULL = 0. They are not zero after
     * ODpair::delete_spurious_routes (delete_bad_routes=1, B/od.cpp:160-244) has priced the routes — pruning and ordering the
     * route sets is the host's job (mezzo_lib's own add_od_routes, or mezzo_b200/flatten.py), the cache state decides
     * which calls of Route::cost return a stale sum later (B/route.cpp:173-197). */
    const double *route_sumcost0, *route_last0;
    /* Demand slices (demand.dat "slices:", Network::readrates B/network.cpp:5544-5606): the rate changes the MatrixActions
     * apply (MatrixAction::execute → ODpair::set_rate → ODaction::set_rate, B/network.cpp:8213-8229, B/od.h:63,102), flattened
     * to one entry per (slice, OD pair) in execution order — by load time, equal times in file order. From slice_time on the
     * pair's ODServer runs with mu = 3600 / slice_rate and the ODaction is active (also with a rate of 0, which books its
     * next event at +inf). A MatrixAction precedes every other event of its own instant (it was booked first). */
    int32_t n_slice_rates;
    const int32_t *slice_od;      /* [n_slice_rates] index into od_rate */
    const double *slice_time, *slice_rate;
    /* out (may be NULL): [2] = MzNet::phantom_final_t / _tp for this demand */
    double *phantom_final;
    /* the transit layer's share of Network::init's clock, between the signal controls and the OD pairs: one 0.00001 step per
     * Busline and per active ODstops pair (B/network.cpp:7672-7694) — the same count as MzNet::n_transit_init_steps; 0
     * without a transit layer. It shifts the first booking of every OD pair. */
    int32_t n_transit_init_steps;
} MzDemand;
int mz_trips_generate(const MzDemand *d, int64_t cap, double *time, double *prev_time, int32_t *od, int32_t *route,
                      int64_t *n_out, int32_t *n_odpairs_initialised);

/* The vehicle class of every generated trip (several vehicle classes in vehiclemix.dat). Replaces the call
 * `odpair->vehtypes()->random_vtype()` of ODaction::execute (B/od.cpp:79): ONE Random stream for the whole network (Vtypes::random,
 * seeded with randseed by Vtypes::initialize, B/vtypes.cpp:36-49), one urandom per executed active ODaction — i.e. per trip, in
 * the global event order mz_trips_generate returns the trips in; the class is the first one whose cumulative probability
 * reaches the draw (B/vtypes.h:61-71; the last class if rounding leaves none, where the reference returns NULL).
 * `probability` = the classes' probabilities as Vtypes::initialize leaves them (normalised by their sum unless it is exactly 1,
 * B/vtypes.cpp:52-64) in the order of Vtypes::vtypes at run time. type_index[i] = position in that order. Host only. */
int mz_vtypes_draw(int64_t randseed, int32_t n_types, const double *probability, int64_t n_trips, int32_t *type_index);

/* Replaces the object graph built by executemaster + Network::init (B/network.cpp:7657-7754): uploads the network,
 * seeds every action's first event exactly like init() does. The call itself is synchronous and single-entry like the rest of
 * the ABI, but inside it the tables are built by short-lived worker threads (up to ~12 at the 10 M-trip size) and the large
 * ones are copied to the device by one of them while the others still work; all of them have ended when it returns, and the
 * caller's buffers are not referenced afterwards. */
int mz_sim_create(int device, const MzNet *net, mz_sim **out);
/* Replaces: while (time<runtime) time=eventlist->next();  (Network::step, B/network.cpp:7804-7807). One call runs the
 * whole horizon; calling it again after mz_sim_reset repeats the run (Network::reset, B/network.cpp:310-419). */
int mz_sim_run(mz_sim *s);
int mz_sim_reset(mz_sim *s);
/* Transit overlay hooks (row f3): the event loop in slices, so that the host's transit model can run beside it.
 *   mz_sim_run_until(t)      executes every event with time < min(t, runtime) and returns; repeated calls with growing t
 *                            continue the same run; mz_sim_run finishes it (rest of the horizon + the event beyond it).
 *   mz_sim_stop_arrivals     drains the stop visits booked since the last call (any order; t_arrive >= the time of the
 *                            event that booked it). Busstop::execute's "bus enters the stop" branch is the host's.
 *   mz_sim_stop_departures   books "bus leaves its stop at t_depart" (Busstop::execute's exit branch, B/busline.cpp:
 *                            1196-1263): at that time the engine evaluates density_running_only and Sdfunc::speed of the
 *                            link, and either books the visit of the trip's next stop if it lies on the same link
 *                            (reported like any other visit) or gives the bus its exit time for the REST of the link and
 *                            inserts it with Q::enter_veh. t_depart must not precede the time the engine has run to, so
 *                            the slice length is bounded by the host model's shortest dwell time (its lookahead).
 * Among events of exactly equal time a departure executes first. Single-GPU engines only. */
int mz_sim_run_until(mz_sim *s, double t);
int mz_sim_stop_arrivals(mz_sim *s, mz_stop_visit *out, int64_t cap, int64_t *n_out);
int mz_sim_stop_departures(mz_sim *s, int64_t n, const int32_t *veh, const double *t_depart);
/* Replaces reading ODpair::grid rows (Vehicle::report, B/vehicle.cpp:79-91): arrival[n_trips] = end_time or -1. */
int mz_sim_get_arrivals(mz_sim *s, double *arrival_time /*[n_trips]*/);
/* Exit log, ordered by (vid, position on route). cap in records; n_out receives the total available. */
int mz_sim_get_exit_log(mz_sim *s, MzExit *out, int64_t cap, int64_t *n_out);
/* Replaces Link::avgtimes (LinkTime per link, B/link.cpp:536-556): avg[n_links*n_periods], NaN where no period was closed. */
int mz_sim_get_linktimes(mz_sim *s, double *avg);
/* Replaces Link::end_of_simulation (B/link.cpp:118-141) followed by reading Link::avgtimes, i.e. the values
 * Link::write_time blends with the historical times (B/link.cpp:741-752): the running mean of the period that is still
 * open is stored, periods without a value fall back to the historical time. The engine's state is not changed. */
int mz_sim_get_linktimes_final(mz_sim *s, double *avg);
int mz_sim_last_stats(const mz_sim *s, mz_sim_stats *out);
int mz_sim_set_stream(mz_sim *s, void *cuda_stream);
/* Tuning knobs (never change results):
 *   MZ_SIM_OPT_MAX_EVENTS   cap on the events one node executes per visit (bounds the latency of a visit)
 *   MZ_SIM_OPT_CHECK_EVERY  accepted and ignored (the run is one persistent kernel, the host never polls)
 *   MZ_SIM_OPT_BLOCKS       thread blocks of the persistent node-visit kernel (0 = one per SM, capped by the node count)
 *   MZ_SIM_OPT_KERNEL       node-visit kernel: 0 = chosen by network size, 1 = a warp per node, nodes found by sweeping,
 *                           2 = a thread per node, 3 = a warp per node, nodes woken by the neighbour that unblocked them
 *                           (same results from all of them) */
enum { MZ_SIM_OPT_MAX_EVENTS = 1, MZ_SIM_OPT_CHECK_EVERY = 2, MZ_SIM_OPT_BLOCKS = 3, MZ_SIM_OPT_KERNEL = 4 };
int mz_sim_set_option(mz_sim *s, int option, int64_t value);
int mz_sim_destroy(mz_sim *s);

/* ---- the link update over several GPUs ----------------------------------------------------------------
 * One process per GPU. Every rank creates the SAME network with its own rank and the common node partition; a node owns
 * its turnings/sinks/origins, a link lives on the rank of its DOWNSTREAM node. The shared mutable state of a rank (link
 * records, queue slab, published node keys) sits in one device allocation, the arena; ranks map each other's arenas
 * (CUDA IPC -> NVLink peer memory) and the kernels read/write the home copy directly: the hand-off of a vehicle across a
 * cut link is a 32-byte peer store, the synchronisation is pairwise through the published node keys. No collective.
 * Call order per rank: mz_sim_create_part; exchange handles (any transport); mz_sim_attach_peer for every other rank;
 * [barrier]; mz_sim_run on all ranks concurrently; [barrier]; mz_sim_min_key on all ranks, the owner of the smallest key
 * gets mz_sim_final_event(node, t), the others mz_sim_final_event(-1, t) (the event beyond the horizon);
 * outputs are per rank (a vehicle's exits are logged by the rank of the exiting node) and merged by the caller
 * (mezzo_b200/dist.py::PartitionedSimulation). */
int mz_sim_create_part(int device, const MzNet *net, int n_ranks, int rank, const int32_t *node_rank /*[n_nodes]*/,
                       mz_sim **out);
int mz_sim_arena_handle(mz_sim *s, void *handle64 /* cudaIpcMemHandle_t */, int64_t *arena_bytes);
int mz_sim_attach_peer(mz_sim *s, int peer_rank, const void *handle64);
/* The same for several GPUs driven by ONE process (host/NetworkStep.h with MEZZO_B200_GPUS=N: one host thread per engine):
 * direct peer access between the devices instead of IPC handles. */
int mz_sim_attach_peer_local(mz_sim *s, int peer_rank, mz_sim *peer);
/* A node partition for mz_sim_create_part: graph-growing recursive bisection balanced by the expected load of the nodes
 * (the METIS initial-partition heuristic; the reference has no counterpart, it is single-threaded). node_rank[n_nodes]. */
int mz_partition_nodes(const MzNet *net, int n_ranks, int32_t *node_rank);
int mz_sim_min_key(mz_sim *s, double *t, double *t_parent, int32_t *sub, int32_t *node /* -1: none */);
int mz_sim_final_event(mz_sim *s, int32_t final_node, double end_time);

#ifdef __cplusplus
}
#endif
#endif /* MEZZO_B200_H */
