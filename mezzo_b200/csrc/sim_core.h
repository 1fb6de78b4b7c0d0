// Hot path 1: the link update — Network::step's event loop recast as an asynchronous conservative simulation: every node
// executes its events that precede its neighbours' published keys; no global window, no grid barrier.
//
// Reference (B/ = /root/reference/branches/busmezzo/mezzo_lib/src/):
//   Network::init / step            B/network.cpp:7657-7754, 7792-7821      Eventlist           B/eventlist.h:58-71
//   TurnAction / Turning            B/turning.cpp:45-146, 237-265           Link                B/link.cpp:155-183, 320-390, 418-655, 670-687
//   Q                               B/q.cpp:52-129, 156-249, 346-361         Sdfunc              B/sdfunc.cpp:12-53
//   Server / Random                 B/server.cpp:32-50, B/Random.cpp:85-132  Daction / Origin    B/node.cpp:126-250, 637-673
//
// Execution model: every node (junction, origin, destination) is a logical process that owns its actions
// (TurnActions of its turnings, Dactions of its in-links, or the trips it generates). An action never wakes another
// action: each one re-books only itself. Therefore a node's earliest pending event key only ever grows and is a hard
// lower bound on everything that node will ever do, so a node may safely execute all of its events that precede the
// published keys of its neighbours (nodes sharing a link with it). This is an ASYNCHRONOUS conservative scheme: there
// is no global window and no global barrier — a node runs as soon as its own key is a local minimum, fences, and
// publishes its new key; the lookahead is per node and adaptive instead of the global min-free-flow-time window because
// spill-back (Link::full), shock-wave unblocking and the running-segment density act with ZERO lookahead.
// Adjacent nodes are never active at the same time (each would need a key below the other's), so link queues (shared
// only by a link's two end nodes) need no atomics and the result equals the sequential event order whatever the timing.
// Ties in time are ordered like the reference's FIFO multimap by (time, time of the booking event, Lamport counter
// inside the instant); full keys are published under a sequence lock and only read when two neighbours tie in time.
//
// Data layout in HBM — packed records, one 32-byte sector per object, so an event is ~6 dependent load rounds:
//   LinkStat (static)   qoff,qcap,maxcap,lanes,length,sdfunc,freeflow        LinkHot (mutable)  q_head,q_size,nr_exits_blocked,blocked_until
//   LinkAcc  (mutable)  travel-time accumulators of Link::exit_veh           QEnt               exit,entry,veh,next,rpos,logi (list order!)
//   ActDyn   (mutable)  t,t_parent,sub                                       ActStat/TurnStat/SrvStat/SdRec (static)
// There are no per-vehicle objects at all: what the reference keeps in Vehicle (entry time, position on the route) rides
// inside the queue entry, so handing a vehicle to the next link — also across GPUs — is one 32-byte store.
// Shared mutable state (LinkHot, queue slab, published node keys) lives in an "arena" with identical layout on every rank.
// Objects on the cut (a link whose end nodes sit on different GPUs, a node with a neighbour on another GPU) are
// REPLICATED: every rank reads only its own arena, and the (single) writer of the moment stores each change into its own
// copy and — through NVLink peer pointers — into the copy of the other rank(s) before it publishes its key. Since the two
// end nodes of a link never run at the same time the copies never diverge; a cut costs asynchronous peer stores and one
// system-scope fence, not a chain of NVLink round trips (first version: every access at the home rank, +14 us per crossing).
//
// This header holds the data layout and every per-event function; it compiles both as CUDA device code (sim.cu, the
// product) and as plain C++ (tests/emu/sim_emu.cpp: a TEST-ONLY host emulation of the window schedule that lets the CPU
// test-suite check the window logic and the multi-rank hand-off without a GPU; the product never loads it).
#pragma once
#include <cmath>
#include <cstddef>
#include <cstdint>
#include <cstring>

#ifdef __CUDACC__
#define MZ_UNROLL _Pragma("unroll")
#define MZ_HD __device__
#define MZ_HDX __host__ __device__
#define MZ_NOINLINE __forceinline__
// Rare or self-contained pieces are real functions: the node-visit kernel had grown to 285 KB of SASS (the double-precision
// log / sin / pow / exp sequences of Server::next and Sdfunc::speed inlined at every call site), far beyond the SM's
// instruction caches, and twelve warps on twelve different paths kept missing them. The kernels take their SimDev as a
// __grid_constant__ parameter, so passing `d` by reference to such a function does not copy it.
#ifdef MZ_COLD_INLINE
#define MZ_COLD
#else
#define MZ_COLD __noinline__
#endif
#define MZ_ATOMIC_ADD(p, v) atomicAdd((p), (unsigned long long)(v))
#else
#define MZ_UNROLL
#define MZ_HD
#define MZ_HDX
#define MZ_NOINLINE
#define MZ_COLD
#define MZ_ATOMIC_ADD(p, v) __atomic_fetch_add((p), (unsigned long long)(v), __ATOMIC_RELAXED)
struct double2 { double x, y; };
static inline double2 make_double2(double x, double y) { double2 r; r.x = x; r.y = y; return r; }
#endif

#ifdef MZ_EMU_TRACE
void mz_emu_trace(double t, int kind, int idx);
#endif

// MZ_GROUP_W (sim_group.cu): the same per-event code compiled for a "warp" of MZ_GROUP_W lanes (1, 2, 4, 8 or 16), i.e.
// several nodes are visited concurrently by the lane groups of one hardware warp — the variant for networks with far
// more nodes than resident warps. It lives in its own namespace so that both variants can be linked into one library.
#ifndef MZ_SIM_NS
#define MZ_SIM_NS mzsim
#endif
#if defined(__CUDACC__) && !defined(MZ_GROUP_W)
#define MZ_WARP_COOP 1
#endif

namespace MZ_SIM_NS {

constexpr int MZ_MAXR = 8;  // ranks (GPUs of one NVSwitch node)

struct alignas(16) QEnt {
    double exit;   // earliest exit time (list is ordered by insertion rule, not strictly by this key — SURVEY H6)
    double entry;  // Vehicle::entry_time on this link
    int veh;       // trip index
    int next;      // Vehicle::nextlink(): route successor of the link the entry sits in, -1 at the route's end
    int rpos;      // index of this link in route_term (route links with a -1 terminator per route)
    int logi;      // slot of this traversal in the exit log
};
static_assert(sizeof(QEnt) == 32, "queue entry must be 32 bytes");
struct alignas(16) LinkStat {
    int qoff, qcap, maxcap, lanes;
    int length, sdfunc;
    double freeflow;
};
static_assert(sizeof(LinkStat) == 32, "LinkStat must be 32 bytes");
struct alignas(16) LinkHot {
    int q_head, q_size;
    int nr_exits_blocked;
    int sd_cur;  // Link::sdfunc while an incident has replaced it: sdfunc index + 1, 0 = the link's own function (LinkStat)
    double blocked_until;
    int sd_tmp, pad;  // Link::temp_sdfunc, same encoding (B/link.cpp:854-868)
};
static_assert(sizeof(LinkHot) == 32, "LinkHot must be 32 bytes");
struct alignas(16) LinkAcc {
    double avg_time, tmp_avg;
    int nr_passed, tmp_passed, curr_period, pad;
};
static_assert(sizeof(LinkAcc) == 32, "LinkAcc must be 32 bytes");
struct alignas(16) ActDyn {
    double t, tp;
    int sub, pad;
    double pad2;
};
static_assert(sizeof(ActDyn) == 32, "ActDyn must be 32 bytes");
struct alignas(16) ActStat {
    int kind;  // ACT_*
    int idx;   // turning | in-link of the sink | OD pair
    int aux;   // sink: slot of its server seed | OD: origin index
    int aux2;  // sink: server index or -1 (the Destination's own default server)
    // a TurnAction carries its turning's record (TurnStat) along: one load instead of the chain action -> turning
    int in, out, server, lookback;
};
static_assert(sizeof(ActStat) == 32, "ActStat must be 32 bytes");
struct alignas(16) TurnStat {
    int in, out, server, lookback;
};
struct alignas(16) SrvStat {
    double mu, sd, delay;
    int type, chg;  // chg: first entry of the server's rate-change list in SimDev::srv_chg, -1 = the rates never change
};
// one ChangeRateAction (B/server.cpp:146-168), already folded: the server's (mu, sd) from time t on; every server's list is
// sorted by time and closed by an entry with t = +inf
struct alignas(16) SrvChg {
    double t, mu, sd, pad;
};
struct alignas(16) SdRec {
    double vfree, vmin, romax, romin, alpha, beta;
    int type, pad;
    double pad2;
};
static_assert(sizeof(SdRec) == 64, "SdRec must be 64 bytes");

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

enum { ACT_TURN = 0, ACT_DEST = 1, ACT_OD = 2, ACT_SIGNAL = 3, ACT_INCIDENT = 4, ACT_STOP = 5 };

// Transit overlay hooks. A bus that Link::enter_veh finds on the link of its next stop is not queued
// (B/link.cpp:489-516): its queue-entry fields wait in a BusAt record, the visit goes to the host as a StopVisit (layout of
// mz_stop_visit), and the departure the host books later is an event of the link's ACT_STOP action.
struct alignas(16) BusAt {
    double entry;  // Vehicle::entry_time on the link
    int rpos, logi;
};
struct alignas(16) StopVisit {
    int veh, stop, link, pad;
    double t_enter, t_arrive;
};
static_assert(sizeof(StopVisit) == 32, "StopVisit must match mz_stop_visit");

// Traffic signals (B/trafficsignal.cpp). One signal EVENT of a control as the host precomputed it (sim_host.h): the turnings
// it stops / activates in order (sig_ops: turning << 1 | activate), the time of the event that booked it (tp) and that
// event's index (parent, -1 = Network::init, whose Lamport counter is sub0).
struct alignas(16) SigEv {
    double t, tp;
    int op_off, n_ops, parent, sub0;
};
static_assert(sizeof(SigEv) == 32, "SigEv must be 32 bytes");
#define MZ_SIG_CHAINS 4  // event-chain slots per TurnAction of a signalised turning

// everything the kernels need, passed by value
struct SimDev {
    int n_nodes, n_links, n_turnings, n_trips, n_periods, n_acts;
    int max_events;  // cap on events one node executes per window (bounds the window's latency; never changes results)
    int n_ranks, rank;
    int server_type;  // Parameters::server_type: 3 / 4 switch every class-Server object to log-normal / log-logistic headways
    double runtime, periodlength;
    // static network (replicated on every rank)
    const LinkStat *lstat;
    const double *link_hist;
    const SdRec *sd;
    const SrvStat *srv;
    const SrvChg *srv_chg;
    const double *inc_stop;   // incidents (null without): end time and replacement sdfunc of each
    const int *inc_sdfunc;
    // traffic signals (all null without signal controls)
    const SigEv *sig_ev;
    const int *sig_ops, *sigc_ev0, *turn_slot0, *turn_nact;
    const unsigned char *node_sig;  // node owns a signal action: its action keys are not cached across events
    unsigned char *turn_active;     // Turning::active
    int *sig_next, *sig_endsub;     // per control: next event (relative); per event: Lamport counter at its end
    const TurnStat *turn;
    const ActStat *astat;
    const int *route_term;  // all routes' links, each route followed by -1
    const int *node_act_off, *node_nb_off;
    const int *node_nb;  // neighbour | its sharing class << MZ_NB_SHIFT
    const int *origin_trip_off, *origin_trips;  // trips per origin in ODaction execution order (virtual-queue order)
    const int *od_off, *od_trips;               // trips per OD pair (one ODaction each), in order
    const int *trip_slot;                       // position of a trip inside its origin's list
    const int *trip_rpos0;                      // route_term index of the trip's first link
    const long long *trip_log_off;
    const double *trip_time, *trip_length;
    const unsigned char *link_home, *node_home;  // owning rank (links: rank of the downstream node)
    const unsigned char *link_peer;              // per rank: the OTHER rank holding a copy of a cut link touching this rank, 255 = none
    const unsigned char *node_peers;             // per rank: bit mask of the other ranks holding a copy of this rank's node state
    // Sharing class of a node's published state (key, last instant) and of a link's state (LinkHot, queue) — who, besides
    // the owner, touches it:  0 = only nodes of the same REGION (one thread block, one SM: the state is read through that
    // SM's L1 and ordered with block-scope fences), 1 = a node of another region of this GPU (read past L1, gpu-scope
    // release), 2 = a node on another rank (read past L1 from this rank's replica, system-scope release).
    const unsigned char *node_cls, *link_cls;
    const int *region_off;                        // my_nodes[region_off[r] .. region_off[r+1]) = the nodes of region r
    int n_regions;
    // Wake-up tables of the warp-per-node kernel (sim.cu): a node that published a new key tells exactly the neighbours it
    // may have unblocked. nb_slot[i] (parallel to node_nb): the neighbour's position inside the SAME region (>= 0: a bit of
    // the block's ready mask in shared memory), -(g + 2) for a neighbour in another region of this rank (g = its bit in the
    // rank's mailbox words, mb_off[region] * 32 + position), -1 for a neighbour on another rank (those nodes are polled).
    const int *nb_slot;
    const int *mb_off;                            // [n_regions + 1] first mailbox word of each region
    unsigned *mailbox;                            // ready bits set by other regions (L2 atomics), drained by the owning block
    const int *poll_off, *poll_pos;               // per region: positions of the nodes with a neighbour on another rank
    const int *my_nodes;                          // nodes of this rank
    int n_my;
    // shared mutable state: per-rank arenas of identical layout, accessed at the object's home rank
    char *arena[MZ_MAXR];
    size_t off_hot, off_slab, off_kt, off_kver, off_kslot[2], off_last;
    // rank-local mutable state (only the owning node ever touches it)
    LinkAcc *lacc;
    double *avgtimes;
    ActDyn *adyn;
    unsigned char *turn_blocked;
    long long *srv_seed;
    int *od_next, *origin_vq_head;
    unsigned char *trip_parked;
    // transit overlay hooks (null without buses): stops per trip (static), per-bus state indexed by the trip's FIRST stop
    // slot, visit records for the host, and the departures the host booked — per stop link, sorted by time, constant
    // during a launch — with the link's cursor into them
    const int *trip_stop_off, *stop_link;
    const double *stop_pos;
    int *bus_next;       // stops the bus has been booked for so far
    BusAt *bus_at;
    StopVisit *stop_out;  // counters[6] records
    const int *dep_off, *dep_veh;
    const double *dep_t;
    int *slink_cur;
    // outputs (rank-local, merged by the host)
    double *arrival;
    double2 *exit_log;  // {entry, exit} per (vehicle, route position)
    // counters: [0] traversals [1] exits [2] events [3] arrived [4] node visits that executed events [5] error flag
    //           [6] stop visits reported [7] buses refused by a full queue after a stop
    unsigned long long *counters;
};

// ------------------------------------------------------------------------------------------------ memory access
// Three kinds of memory, three access paths (measured on B200: L1 hit ~32 cycles, L2 hit ~250; __threadfence() compiles to
// MEMBAR.SC.GPU + CCTL.IVALL, i.e. it throws the SM's whole L1 away, fence.release.gpu to MEMBAR.ALL.GPU alone):
//  * static tables — read-only path (ld.global.nc), L1-resident for the whole persistent kernel;
//  * state only ONE SM ever touches (a node's own actions, seeds, accumulators; key / last instant / link state / queue of
//    objects whose users all sit in one region = one thread block) — ordinary loads and stores through that SM's L1,
//    ordered between the warps of the block by __threadfence_block(), which does not invalidate anything;
//  * state shared between SMs or GPUs — strong loads past L1 (ld.global.cg) and a RELEASE fence (no L1 invalidation)
//    before the key that publishes it. No acquire fence is needed: nothing that another SM writes is ever read through L1.
constexpr int MZ_NB_SHIFT = 30;
constexpr int MZ_NB_MASK = (1 << MZ_NB_SHIFT) - 1;
#ifdef __CUDACC__
MZ_HD inline int ldm(const int *p) { return __ldcg(p); }
MZ_HD inline double ldm(const double *p) { return __ldcg(p); }
MZ_HD inline long long ldm(const long long *p) { return __ldcg(p); }
MZ_HD inline unsigned char ldm(const unsigned char *p) { return __ldcg(p); }
MZ_HD inline unsigned ldm(const unsigned *p) { return __ldcg(p); }
template <class T>
MZ_HD inline T ldm32(const T *p)  // 32-byte record
{
    static_assert(sizeof(T) == 32, "ldm32 needs a 32-byte record");
    union {
        T v;
        int4 q[2];
    } u;
    u.q[0] = __ldcg(reinterpret_cast<const int4 *>(p));
    u.q[1] = __ldcg(reinterpret_cast<const int4 *>(p) + 1);
    return u.v;
}
// single-SM state: ordinary (weak) loads through L1
template <class T>
MZ_HD inline T ldp(const T *p) { return *p; }
MZ_HD inline int4 ldp16(const void *p) { return *reinterpret_cast<const int4 *>(p); }
template <class T>
MZ_HD inline T ldp32(const T *p)
{
    static_assert(sizeof(T) == 32, "ldp32 needs a 32-byte record");
    union {
        T v;
        int4 q[2];
    } u;
    u.q[0] = ldp16(p);
    u.q[1] = ldp16(reinterpret_cast<const int4 *>(p) + 1);
    return u.v;
}
// The words nodes poll each other through (key time, publication counter) are STRONG accesses at the scope of the
// object's sharing class — ptxas may neither drop nor hoist them, and together with the fences below they form the PTX
// memory model's fence-fence message passing. ld.relaxed.cta is served by L1 (measured: 32 cycles), .gpu/.sys by L2.
MZ_HD inline double ld_flag(const double *p, int cls)
{
    double v;
    if (cls == 0) asm volatile("ld.relaxed.cta.global.f64 %0, [%1];" : "=d"(v) : "l"(p) : "memory");
    else if (cls == 1) asm volatile("ld.relaxed.gpu.global.f64 %0, [%1];" : "=d"(v) : "l"(p) : "memory");
    else asm volatile("ld.relaxed.sys.global.f64 %0, [%1];" : "=d"(v) : "l"(p) : "memory");
    return v;
}
MZ_HD inline unsigned ld_flag(const unsigned *p, int cls)
{
    unsigned v;
    if (cls == 0) asm volatile("ld.relaxed.cta.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    else if (cls == 1) asm volatile("ld.relaxed.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    else asm volatile("ld.relaxed.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
MZ_HD inline void st_pub(double *p, double v, int cls)
{
    if (cls == 0) asm volatile("st.relaxed.cta.global.f64 [%0], %1;" ::"l"(p), "d"(v) : "memory");
    else if (cls == 1) asm volatile("st.relaxed.gpu.global.f64 [%0], %1;" ::"l"(p), "d"(v) : "memory");
    else asm volatile("st.relaxed.sys.global.f64 [%0], %1;" ::"l"(p), "d"(v) : "memory");
}
MZ_HD inline void st_pub(unsigned *p, unsigned v, int cls)
{
    if (cls == 0) asm volatile("st.relaxed.cta.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
    else if (cls == 1) asm volatile("st.relaxed.gpu.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
    else asm volatile("st.relaxed.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
MZ_HD inline void mz_prefetch(const void *p) { asm volatile("prefetch.global.L1 [%0];" ::"l"(p)); }
// possibly shared state: sh = sharing class (0 = single SM)
template <class T>
MZ_HD inline T ldx(const T *p, int sh) { return sh ? ldm(p) : ldp(p); }
template <class T>
MZ_HD inline T ldx32(const T *p, int sh) { return sh ? ldm32(p) : ldp32(p); }
template <class T>
MZ_HD inline T lds(const T *p)  // static record of 16/32/64 bytes
{
    union {
        T v;
        int4 q[sizeof(T) / 16];
    } u;
MZ_UNROLL
    for (unsigned i = 0; i < sizeof(T) / 16; ++i) u.q[i] = __ldg(reinterpret_cast<const int4 *>(p) + i);
    return u.v;
}
MZ_HD inline int lds(const int *p) { return __ldg(p); }
MZ_HD inline double lds(const double *p) { return __ldg(p); }
MZ_HD inline long long lds(const long long *p) { return __ldg(p); }
MZ_HD inline unsigned char lds(const unsigned char *p) { return __ldg(p); }
// Full fence of the given sharing class — only the (rare) full-key reads use it.
MZ_HD inline void mz_fence(int cls)
{
    if (cls == 0) __threadfence_block();
    else if (cls == 1) __threadfence();
    else __threadfence_system();
}
// Release: everything this thread wrote (and what the other lanes wrote before the last __syncwarp) becomes visible to
// the readers of the given class before anything it writes afterwards. No L1 invalidation.
MZ_HD inline void mz_release(int cls)
{
    if (cls == 0) __threadfence_block();
    else if (cls == 1) asm volatile("fence.release.gpu;" ::: "memory");
    else asm volatile("fence.release.sys;" ::: "memory");
}
// Acquire side of a visit: the neighbours' keys were read with strong loads; what they published is read past L1 (classes
// 1, 2) or lives in this SM's L1 (class 0). A block-scope fence (MEMBAR.CTA, ~20 cycles, no L1 invalidation) keeps both
// the compiler and the hardware from starting those reads before the keys have arrived; for class 0 it is the formal
// acquire fence of the pattern, for classes 1-2 a gpu/system-scope acquire would only add the invalidation of an L1 that
// holds nothing another SM writes.
MZ_HD inline void mz_acquire(int) { __threadfence_block(); }
#else
// the multi-rank emulation maps the arenas into several processes: loads of mutable words must really be performed
template <class T> inline T ldm(const T *p) { return *const_cast<const volatile T *>(p); }
template <class T> inline T ldm32(const T *p) { T v; std::memcpy(&v, const_cast<const T *>(p), sizeof(T)); return v; }
template <class T> inline T ldp(const T *p) { return ldm(p); }
template <class T> inline T ldp32(const T *p) { return ldm32(p); }
template <class T> inline T ldx(const T *p, int) { return ldm(p); }
template <class T> inline T ldx32(const T *p, int) { return ldm32(p); }
template <class T> inline T lds(const T *p) { return *p; }
inline void mz_fence(int) { __atomic_thread_fence(__ATOMIC_SEQ_CST); }
inline void mz_release(int) { __atomic_thread_fence(__ATOMIC_SEQ_CST); }
inline void mz_acquire(int) { __atomic_thread_fence(__ATOMIC_SEQ_CST); }
template <class T> inline T ld_flag(const T *p, int) { return ldm(p); }
inline void mz_prefetch(const void *) {}
template <class T> inline void st_pub(T *p, T v, int) { *const_cast<volatile T *>(p) = v; }
#endif

// ------------------------------------------------------------------------------------------------ Random / Server
MZ_HD inline double urandom(long long &seed)  // B/Random.cpp:85-98
{
    // the state stays in [1, M-1] and both products below fit 32 bits (A*(Q-1) < 2^31, R*(M/Q) < 2^28): plain int arithmetic
    // gives the very same integers as the reference's long arithmetic at a third of the instructions
    if ((unsigned long long)seed > 2147483647ULL) {  // a user seed beyond 31 bits (first draw only): the reference's long arithmetic
        const long long M = 2147483647LL, A = 48271LL, Q = M / A, R = M % A;
        long long s = A * (seed % Q) - R * (seed / Q);
        s = (s > 0) ? s : (s + M);
        seed = s;
        return (double)s / (double)M;
    }
    const int M = 2147483647, A = 48271, Q = M / A, R = M % A;
    const int x = (int)seed;
    int s = A * (x % Q) - R * (x / Q);
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
MZ_HD inline double lnrandom(long long &seed, double m, double v)  // B/Random.cpp:151-157
{
    const double mu_ = log(m * m / sqrt((m * m) + (v * v)));
    const double sd_ = sqrt(log(1 + (v * v) / (m * m)));
    return exp(nrandom(seed, mu_, sd_));
}
MZ_HD inline double loglogisticrandom(long long &seed, double alpha, double beta)  // B/Random.cpp:165-172
{
    const double alpha_logistic = log(alpha);
    const double beta_logistic = 1 / beta;
    // "log(urandom()/(1-urandom()))": the compiled reference draws the numerator's number first (checked bit for bit
    // against oracle/_ref by the loglogistic golden case)
    const double u_num = urandom(seed);
    const double u_den = urandom(seed);
    const double logistic = alpha_logistic + beta_logistic * log(u_num / (1 - u_den));
    return exp(logistic);
}
// type = the server's class (net.dat), gtype = Parameters::server_type, which switches the distribution of every
// class-Server object — type 1 and the destinations' own default servers (B/server.cpp:32-50)
struct ServerDraw {
    double t;
    long long seed;
};
MZ_HD inline double server_next_impl(int type, int gtype, double mu, double sd, double delay, long long &seed, double t)
{
    if (type == 0) return t;                                       // ConstServer            B/server.h:74
    if (type == 2) return t + mu + delay;                          // DetServer              B/server.h:98
    if (type == 4) return t + loglogisticrandom(seed, mu, sd);     // LogLogisticDelayServer B/server.cpp:139-142
    if (gtype == 3) return t + (mu + lnrandom(seed, delay, sd));   // Server::next           B/server.cpp:34-40
    if (gtype == 4) return t + loglogisticrandom(seed, mu, sd);    //                        B/server.cpp:42-45
    const double x = nrandom(seed, mu, sd);                        //                        B/server.cpp:46-48
    return t + (x > 0.1 ? x : 0.1);
}
// (one out-of-line copy of the distributions per kernel; the seed travels by value so that it stays in registers)
MZ_HD MZ_COLD inline ServerDraw server_draw(int type, int gtype, double mu, double sd, double delay, long long seed, double t)
{
    ServerDraw r;
    r.t = server_next_impl(type, gtype, mu, sd, delay, seed, t);
    r.seed = seed;
    return r;
}
MZ_HD inline double server_next(int type, int gtype, double mu, double sd, double delay, long long &seed, double t)
{
    if (type == 0) return t;               // ConstServer / DetServer need no call
    if (type == 2) return t + mu + delay;
    const ServerDraw r = server_draw(type, gtype, mu, sd, delay, seed, t);
    seed = r.seed;
    return r.t;
}
// the server's (mu, sd) as seen by an event at time t: every ChangeRateAction with time <= t has executed (it was booked
// before all actions, so it also precedes the events of its own instant)
template <class Dev>
MZ_HD inline void server_rates_at(const Dev &d, const SrvStat &sv, double t, double &mu, double &sd)
{
    mu = sv.mu;
    sd = sv.sd;
    if (sv.chg >= 0)
        for (int i = sv.chg;; ++i) {
            const SrvChg c = lds(d.srv_chg + i);
            if (!(c.t <= t)) break;
            mu = c.mu;
            sd = c.sd;
        }
}
MZ_HD MZ_COLD inline double sd_speed_dynamit(double vmin, double vfree, double x, double alpha, double beta)  // B/sdfunc.cpp:45-53
{
    return vmin + (vfree - vmin) * pow((1 - pow(x, alpha)), beta);
}
MZ_HD inline double sd_speed(const SdRec &f, double ro)  // B/sdfunc.cpp:12-19, 45-53
{
    if (ro <= f.romin) return f.vfree;
    if (ro > f.romax) return f.vmin;
    const double x = (ro - f.romin) / (f.romax - f.romin);
    if (f.type == 0) return f.vmin + (f.vfree - f.vmin) * (1 - x);
    return sd_speed_dynamit(f.vmin, f.vfree, x, f.alpha, f.beta);
}

// ------------------------------------------------------------------------------------------------ warp abstraction
// One WARP cooperates on one node: every lane runs the same control flow on the same (uniform) scalars, queue scans
// and shifts are spread over the lanes (ballot to locate), and lane 0 alone stores.
// The host emulation compiles the same code with a "warp" of one lane, the lane-group device variant (MZ_GROUP_W) with
// a "warp" of 1-16 lanes.
#ifdef MZ_WARP_COOP
#define MZ_W 32
MZ_HD inline int mz_lane() { return (int)(threadIdx.x & 31u); }
MZ_HD inline unsigned mz_ballot(bool p) { return __ballot_sync(0xffffffffu, p); }
MZ_HD inline void mz_syncwarp() { __syncwarp(); }
MZ_HD inline int mz_ffs(unsigned b) { return __ffs((int)b); }
MZ_HD inline int mz_clz(unsigned b) { return __clz((int)b); }
MZ_HD inline double mz_shfl_xor(double v, int m) { return __shfl_xor_sync(0xffffffffu, v, m); }
MZ_HD inline int mz_shfl_xor(int v, int m) { return __shfl_xor_sync(0xffffffffu, v, m); }
MZ_HD inline double mz_shfl(double v, int src) { return __shfl_sync(0xffffffffu, v, src); }
MZ_HD inline int mz_shfl(int v, int src) { return __shfl_sync(0xffffffffu, v, src); }
#elif defined(__CUDACC__) && MZ_GROUP_W > 1
// a group of MZ_GROUP_W consecutive lanes of a hardware warp; every primitive names only the group's lanes in its mask,
// so the groups of one warp run independently of each other (independent thread scheduling)
#define MZ_W MZ_GROUP_W
MZ_HD inline unsigned mz_gshift() { return (threadIdx.x & 31u) & ~(unsigned)(MZ_W - 1); }
MZ_HD inline unsigned mz_gmask() { return ((1u << MZ_W) - 1u) << mz_gshift(); }
MZ_HD inline int mz_lane() { return (int)(threadIdx.x & (unsigned)(MZ_W - 1)); }
MZ_HD inline unsigned mz_ballot(bool p) { return (__ballot_sync(mz_gmask(), p) >> mz_gshift()) & ((1u << MZ_W) - 1u); }
MZ_HD inline void mz_syncwarp() { __syncwarp(mz_gmask()); }
MZ_HD inline int mz_ffs(unsigned b) { return __ffs((int)b); }
MZ_HD inline int mz_clz(unsigned b) { return __clz((int)b); }
MZ_HD inline double mz_shfl_xor(double v, int m) { return __shfl_xor_sync(mz_gmask(), v, m, MZ_W); }
MZ_HD inline int mz_shfl_xor(int v, int m) { return __shfl_xor_sync(mz_gmask(), v, m, MZ_W); }
MZ_HD inline double mz_shfl(double v, int src) { return __shfl_sync(mz_gmask(), v, src, MZ_W); }
MZ_HD inline int mz_shfl(int v, int src) { return __shfl_sync(mz_gmask(), v, src, MZ_W); }
#else
#define MZ_W 1
MZ_HD inline int mz_lane() { return 0; }
MZ_HD inline unsigned mz_ballot(bool p) { return p ? 1u : 0u; }
MZ_HD inline void mz_syncwarp() {}
#ifdef __CUDACC__
MZ_HD inline int mz_ffs(unsigned b) { return __ffs((int)b); }
MZ_HD inline int mz_clz(unsigned b) { return __clz((int)b); }
#else
inline int mz_ffs(unsigned b) { return __builtin_ffs((int)b); }
inline int mz_clz(unsigned b) { return __builtin_clz(b); }
#endif
MZ_HD inline double mz_shfl_xor(double v, int) { return v; }
MZ_HD inline int mz_shfl_xor(int v, int) { return v; }
MZ_HD inline double mz_shfl(double v, int) { return v; }
MZ_HD inline int mz_shfl(int v, int) { return v; }
#endif
// List positions a lane handles per load round: a lane can issue the loads of MZ_U positions back to back before it looks
// at the first one. Measured on B200 with the thread-per-node kernel (1.02 M links): MZ_U = 4 HALVED the throughput
// (5.07 s vs 2.43 s per horizon — the per-lane arrays go to local memory and the 4-wide rounds lengthen the divergent
// sections of a warp), so every build uses 1; the loops below are written for any value.
constexpr int MZ_U = 1;
template <class T>
MZ_HD inline void wr(T *p, T v)  // uniform store: lane 0 only
{
    if (mz_lane() == 0) *p = v;
}
// warp-wide minimum of a key (every lane ends up with the same result). Times are >= 0 (or +inf), so their bit patterns
// order like integers and two 32-bit warp reductions find the smallest time; almost always a single lane holds it and
// the remaining fields are simply broadcast from that lane. Ties fall back to the full lexicographic butterfly.
MZ_HD MZ_COLD inline Key key_warp_min_ties(Key k)  // the full lexicographic butterfly (several lanes share the smallest time)
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
MZ_HD inline Key key_warp_min(Key k)
{
#ifdef MZ_WARP_COOP
    const unsigned long long tb = (unsigned long long)__double_as_longlong(k.t);
    const unsigned hi = (unsigned)(tb >> 32), lo = (unsigned)tb;
    const unsigned mh = __reduce_min_sync(0xffffffffu, hi);
    const unsigned ml = __reduce_min_sync(0xffffffffu, hi == mh ? lo : 0xffffffffu);
    const unsigned eq = __ballot_sync(0xffffffffu, hi == mh && lo == ml);
    if ((eq & (eq - 1)) == 0) {
        const int src = __ffs((int)eq) - 1;
        Key r;
        r.t = __longlong_as_double((long long)(((unsigned long long)mh << 32) | ml));
        r.tp = __shfl_sync(0xffffffffu, k.tp, src);
        r.sub = __shfl_sync(0xffffffffu, k.sub, src);
        r.id = __shfl_sync(0xffffffffu, k.id, src);
        return r;
    }
#endif
    return key_warp_min_ties(k);
}

// warp-wide maximum of (instant, sub-counter), lexicographic; instants are >= 0, sub-counters >= 0
MZ_HD inline void warp_max_instant(double &t, int &sub)
{
#ifdef MZ_WARP_COOP
    const unsigned long long tb = (unsigned long long)__double_as_longlong(t);
    const unsigned hi = (unsigned)(tb >> 32), lo = (unsigned)tb;
    const unsigned mh = __reduce_max_sync(0xffffffffu, hi);
    const unsigned ml = __reduce_max_sync(0xffffffffu, hi == mh ? lo : 0u);
    const unsigned ms = __reduce_max_sync(0xffffffffu, (hi == mh && lo == ml) ? (unsigned)sub : 0u);
    t = __longlong_as_double((long long)(((unsigned long long)mh << 32) | ml));
    sub = (int)ms;
#else
    for (int m = MZ_W / 2; m > 0; m >>= 1) {
        const double ot = mz_shfl_xor(t, m);
        const int os = mz_shfl_xor(sub, m);
        if (ot > t || (ot == t && os > sub)) {
            t = ot;
            sub = os;
        }
    }
#endif
}

// ------------------------------------------------------------------------------------------------ shared-state accessors
// every READ of shared state goes to this rank's own arena
MZ_HD inline char *my_arena(const SimDev &d) { return d.arena[d.rank]; }
MZ_HD inline int link_peer_rank(const SimDev &d, int l)
{
    if (d.n_ranks == 1) return -1;
    const int p = (int)lds(d.link_peer + l);
    return p == 255 ? -1 : p;
}
MZ_HD inline unsigned node_peer_mask(const SimDev &d, int n) { return d.n_ranks > 1 ? (unsigned)lds(d.node_peers + n) : 0u; }
// Published key of node m: the time alone is one 8-byte word (always consistent, and all the common case needs) ...
MZ_HD inline double node_key_time(const SimDev &d, int m, int cls)
{
    return ld_flag(reinterpret_cast<const double *>(my_arena(d) + d.off_kt) + m, cls);
}
MZ_HD inline int node_class(const SimDev &d, int n) { return (int)lds(d.node_cls + n); }
// ... the full key (time, parent time, Lamport counter) is published in one of two 32-byte slots, selected by the parity of
// the node's publication counter `ver`: publication k writes slot[k&1] BEFORE its fence and the time word + ver = k after
// it, so one fence per publication suffices and a reader that sees the same ver before and after reading slot[ver&1] holds
// a consistent key. Any published key — also a slightly older one — is a valid lower bound of the node's future.
struct alignas(16) KeySlot {
    double t, tp;
    int sub, pad;
    double pad2;
};
static_assert(sizeof(KeySlot) == 32, "KeySlot must be 32 bytes");
MZ_HD MZ_COLD inline Key node_key_full(const SimDev &d, int m, int cls)
{
    const char *a = my_arena(d);
    const unsigned *ver = reinterpret_cast<const unsigned *>(a + d.off_kver) + m;
    Key k;
    for (;;) {
        const unsigned v1 = ld_flag(ver, cls);
        mz_fence(cls);
        const KeySlot ks = ldx32(reinterpret_cast<const KeySlot *>(a + d.off_kslot[v1 & 1u]) + m, cls);
        mz_fence(cls);
        const unsigned v2 = ld_flag(ver, cls);
        k.t = ks.t;
        k.tp = ks.tp;
        k.sub = ks.sub;
        if (v1 == v2) break;
    }
    k.id = m;
    return k;
}
// Owner only (lane 0 stores): everything the node wrote during its visit becomes visible before the new key does.
// The key goes into this rank's arena and into the arenas of the ranks that hold a copy of the node (peer stores).
// `ver` = the node's current publication counter (the owner keeps it in its NodeLast record, read at the start of the visit).
MZ_HD inline void node_key_publish(const SimDev &d, int n, const Key &k, unsigned ver, int cls)
{
    if (mz_lane() == 0) {
        const unsigned v = ver + 1u;
        KeySlot ks;
        ks.t = k.t;
        ks.tp = k.tp;
        ks.sub = k.sub;
        ks.pad = 0;
        ks.pad2 = 0.0;
        const unsigned peers = node_peer_mask(d, n) | (1u << d.rank);
        for (unsigned m = peers; m; m &= m - 1) reinterpret_cast<KeySlot *>(d.arena[mz_ffs(m) - 1] + d.off_kslot[v & 1u])[n] = ks;
        mz_release(cls);
        for (unsigned m = peers; m; m &= m - 1) {
            char *a = d.arena[mz_ffs(m) - 1];
            st_pub(reinterpret_cast<double *>(a + d.off_kt) + n, k.t, cls);
            st_pub(reinterpret_cast<unsigned *>(a + d.off_kver) + n, v, cls);
        }
    }
    mz_syncwarp();
}
// (time instant of the node's most recent execution, largest Lamport sub-counter used there): one 16-byte record, read
// with one load
struct alignas(16) NodeLast {
    double t;  // 0 = never (all event times are >= 0.1)
    int sub;
    unsigned ver;  // the node's publication counter (a copy for its owner: saves the owner a load of the key version word)
};
MZ_HD inline NodeLast node_last(const SimDev &d, int m, int cls)
{
    const NodeLast *p = reinterpret_cast<const NodeLast *>(my_arena(d) + d.off_last) + m;
#ifdef __CUDACC__
    union {
        NodeLast v;
        int4 q;
    } u;
    u.q = cls ? __ldcg(reinterpret_cast<const int4 *>(p)) : ldp16(p);
    return u.v;
#else
    (void)cls;
    NodeLast v;
    std::memcpy(&v, const_cast<const NodeLast *>(p), sizeof(v));
    return v;
#endif
}
// owner only: the node's last executed instant, into every copy
MZ_HD inline void node_last_store(const SimDev &d, int n, double t, int sub, unsigned ver)
{
    if (mz_lane() == 0) {
        NodeLast v;
        v.t = t;
        v.sub = sub;
        v.ver = ver;
        for (unsigned m = node_peer_mask(d, n) | (1u << d.rank); m; m &= m - 1)
            reinterpret_cast<NodeLast *>(d.arena[mz_ffs(m) - 1] + d.off_last)[n] = v;
    }
}

// ------------------------------------------------------------------------------------------------ queue (ring buffer in list order)
struct QRef {
    QEnt *base;
    LinkHot *hot;
    QEnt *base2;   // the other rank's copy of a cut link (written through), null otherwise
    LinkHot *hot2;
    int cap, head, size;
    int nr_exits_blocked;
    int sd_cur, sd_tmp;  // incident override of the link's sdfunc (LinkHot)
    int sh;  // sharing class of the link (SimDev::link_cls): 0 = both end nodes in this region, read through L1
    double blocked_until;
};
MZ_HD inline QRef q_open(const SimDev &d, int l, const LinkStat &ls)
{
    QRef q;
    q.hot = reinterpret_cast<LinkHot *>(my_arena(d) + d.off_hot) + l;
    q.base = reinterpret_cast<QEnt *>(my_arena(d) + d.off_slab) + ls.qoff;
    const int peer = link_peer_rank(d, l);
    q.hot2 = peer >= 0 ? reinterpret_cast<LinkHot *>(d.arena[peer] + d.off_hot) + l : nullptr;
    q.base2 = peer >= 0 ? reinterpret_cast<QEnt *>(d.arena[peer] + d.off_slab) + ls.qoff : nullptr;
    q.cap = ls.qcap;
    q.sh = (int)lds(d.link_cls + l);
    const LinkHot h = ldx32(q.hot, q.sh);
    q.head = h.q_head;
    q.size = h.q_size;
    q.nr_exits_blocked = h.nr_exits_blocked;
    q.blocked_until = h.blocked_until;
    q.sd_cur = h.sd_cur;
    q.sd_tmp = h.sd_tmp;
    return q;
}
MZ_HD inline void q_close(QRef &q)  // head and size share one 8-byte word
{
    // An empty ring re-anchors at slot 0: in free-flowing traffic queues run empty all the time, so the live entries of
    // every link stay in the first sectors of its slab instead of wandering through all of it (cache footprint).
    if (q.size == 0) q.head = 0;
    if (mz_lane() == 0) {
        q.hot->q_head = q.head;
        q.hot->q_size = q.size;
        if (q.hot2) {
            q.hot2->q_head = q.head;
            q.hot2->q_size = q.size;
        }
    }
}
MZ_HD inline QEnt *q_at(const QRef &q, int i)
{
    int p = q.head + i;
    if (p >= q.cap) p -= q.cap;
    return q.base + p;
}
// every write of queue / link state goes to both copies of a cut link
MZ_HD inline void q_store(const QRef &q, int i, const QEnt &e)
{
    int p = q.head + i;
    if (p >= q.cap) p -= q.cap;
    q.base[p] = e;
    if (q.base2) q.base2[p] = e;
}
MZ_HD inline void q_store_exit(const QRef &q, int i, double exit)
{
    int p = q.head + i;
    if (p >= q.cap) p -= q.cap;
    q.base[p].exit = exit;
    if (q.base2) q.base2[p].exit = exit;
}
MZ_HD inline void hot_store_blocked_until(const QRef &q, double v)
{
    if (mz_lane() == 0) {
        q.hot->blocked_until = v;
        if (q.hot2) q.hot2->blocked_until = v;
    }
}
// the link's speed-density function right now: its own, unless an incident has replaced it
MZ_HD inline int link_sdfunc(const QRef &q, const LinkStat &ls) { return q.sd_cur ? q.sd_cur - 1 : ls.sdfunc; }
MZ_HD inline void hot_store_sdfunc(const QRef &q, int cur, int tmp)
{
    if (mz_lane() == 0) {
        q.hot->sd_cur = cur;
        q.hot->sd_tmp = tmp;
        if (q.hot2) {
            q.hot2->sd_cur = cur;
            q.hot2->sd_tmp = tmp;
        }
    }
}
MZ_HD inline void hot_store_nr_exits_blocked(const QRef &q, int v)
{
    if (mz_lane() == 0) {
        q.hot->nr_exits_blocked = v;
        if (q.hot2) q.hot2->nr_exits_blocked = v;
    }
}
// smallest list position whose entry heads for link `next`, or size
MZ_HD inline int q_find_first_next(const QRef &q, int next)
{
    for (int base = 0; base < q.size; base += MZ_W * MZ_U) {
        int nx[MZ_U];
MZ_UNROLL
        for (int u = 0; u < MZ_U; ++u) {
            const int i = base + u * MZ_W + mz_lane();
            nx[u] = i < q.size ? ldx(&q_at(q, i)->next, q.sh) : ~next;
        }
        unsigned b = 0;
MZ_UNROLL
        for (int u = 0; u < MZ_U; ++u) b |= mz_ballot(nx[u] == next) << (u * MZ_W);
        if (b) return base + mz_ffs(b) - 1;
    }
    return q.size;
}
// smallest list position whose exit_time > t, or size  (Q::nr_running = size - this, B/q.h:67-68)
MZ_HD inline int q_find_first_running(const QRef &q, double t)
{
    for (int base = 0; base < q.size; base += MZ_W * MZ_U) {
        double ex[MZ_U];
MZ_UNROLL
        for (int u = 0; u < MZ_U; ++u) {
            const int i = base + u * MZ_W + mz_lane();
            ex[u] = i < q.size ? ldx(&q_at(q, i)->exit, q.sh) : -INFINITY;
        }
        unsigned b = 0;
MZ_UNROLL
        for (int u = 0; u < MZ_U; ++u) b |= mz_ballot(ex[u] > t) << (u * MZ_W);
        if (b) return base + mz_ffs(b) - 1;
    }
    return q.size;
}
// largest list position whose exit_time is (gt ? > t : <= t), or -1
MZ_HD inline int q_find_last(const QRef &q, double t, bool gt)
{
    constexpr int C = MZ_W * MZ_U;
    for (int base = ((q.size - 1) / C) * C; base >= 0 && q.size > 0; base -= C) {
        double ex[MZ_U];
MZ_UNROLL
        for (int u = 0; u < MZ_U; ++u) {
            const int i = base + u * MZ_W + mz_lane();
            ex[u] = i < q.size ? ldx(&q_at(q, i)->exit, q.sh) : NAN;  // NaN: neither > t nor <= t
        }
        unsigned b = 0;
MZ_UNROLL
        for (int u = 0; u < MZ_U; ++u) b |= mz_ballot(gt ? (ex[u] > t) : (ex[u] <= t)) << (u * MZ_W);
        if (b) return base + (31 - mz_clz(b));
    }
    return -1;
}
// entries [from, to) move one slot towards the tail (to [from+1, to+1)); chunks run from the tail so nothing unread is hit
MZ_HD inline void q_shift_up(const QRef &q, int from, int to)
{
    mz_syncwarp();
    for (int hi = to; hi > from; hi -= MZ_W * MZ_U) {
        QEnt tmp[MZ_U];
MZ_UNROLL
        for (int u = 0; u < MZ_U; ++u) {
            const int i = hi - 1 - u * MZ_W - mz_lane();
            if (i >= from) tmp[u] = ldx32(q_at(q, i), q.sh);
        }
        mz_syncwarp();
MZ_UNROLL
        for (int u = 0; u < MZ_U; ++u) {
            const int i = hi - 1 - u * MZ_W - mz_lane();
            if (i >= from) q_store(q, i + 1, tmp[u]);
        }
        mz_syncwarp();
    }
}
// entries [from+1, to+1) move one slot towards the head (to [from, to))
MZ_HD inline void q_shift_down(const QRef &q, int from, int to)
{
    mz_syncwarp();
    for (int lo = from; lo < to; lo += MZ_W * MZ_U) {
        QEnt tmp[MZ_U];
MZ_UNROLL
        for (int u = 0; u < MZ_U; ++u) {
            const int i = lo + u * MZ_W + mz_lane();
            if (i < to) tmp[u] = ldx32(q_at(q, i + 1), q.sh);
        }
        mz_syncwarp();
MZ_UNROLL
        for (int u = 0; u < MZ_U; ++u) {
            const int i = lo + u * MZ_W + mz_lane();
            if (i < to) q_store(q, i, tmp[u]);
        }
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
    if (mz_lane() == 0) q_store(q, p, e);
    mz_syncwarp();
}

// ------------------------------------------------------------------------------------------------ Link
// Link::full(time) B/link.cpp:155-183 on an opened out-link
MZ_HD inline bool link_full_t(const QRef &q, const LinkStat &ls, double t)
{
    const double bu = q.blocked_until;
    if (bu == -2.0) return true;
    if (q.size >= ls.maxcap) {
        if (q.nr_exits_blocked > 0) hot_store_blocked_until(q, -2.0);
        return true;
    }
    if (bu == -1.0) return false;
    if (bu <= t) {
        hot_store_blocked_until(q, -1.0);
        return false;
    }
    return true;
}

// Link::enter_veh (B/link.cpp:418-523) = density_running_only (670-687) + Sdfunc::speed + Q::enter_veh (B/q.cpp:52-72).
// `rpos`/`logi` are the vehicle's route position / log slot ON THIS LINK.
MZ_HD MZ_COLD inline void link_enter_veh_generic(const SimDev &d, QRef &q, const LinkStat &ls, int v, int rpos, int logi, double t,
                                         unsigned long long &n_trav)
{
    mz_syncwarp();
    const int next = lds(d.route_term + rpos + 1);  // independent of the scans below: issued first
    // nr_running: elements from the first one with exit_time > t to the end (B/q.h:67-68)
    const int nr = q.size - q_find_first_running(q, t);
    // calc_space: vehicle lengths from the back while exit_time <= t (B/q.cpp:346-361); 0 unless the tail is queued
    // The lanes fetch the queued tail's vehicle lengths together (two load rounds per 32 entries); the sum itself runs
    // in the reference's order — back to front — over registers, so it is bit-identical for any mix of lengths.
    double space = 0.0;
    if (q.size > 0) {
        const int last_running = q_find_last(q, t, true);
        for (int hi = q.size; hi > last_running + 1; hi -= MZ_W * MZ_U) {
            int veh[MZ_U];
            double len[MZ_U];
MZ_UNROLL
            for (int u = 0; u < MZ_U; ++u) {
                const int i = hi - 1 - u * MZ_W - mz_lane();
                veh[u] = i > last_running ? ldx(&q_at(q, i)->veh, q.sh) : -1;
            }
MZ_UNROLL
            for (int u = 0; u < MZ_U; ++u) len[u] = veh[u] >= 0 ? lds(d.trip_length + veh[u]) : 0.0;
MZ_UNROLL
            for (int u = 0; u < MZ_U; ++u) {
                const int left = hi - 1 - u * MZ_W - last_running;  // entries of this sub-chunk behind the running part
                const int cnt = left < MZ_W ? left : MZ_W;
                for (int k = 0; k < cnt; ++k) space += mz_shfl(len[u], k);
            }
        }
    }
    const double queue_space = space / ls.lanes;
    const double running_length = ls.length - queue_space;
    double ro = 0.0;
    if (nr && running_length) ro = nr / (running_length * (ls.lanes / 1000.0));
    const double speed = sd_speed(lds(d.sd + link_sdfunc(q, ls)), ro);
    QEnt e;
    e.exit = t + (ls.length / speed);
    e.entry = t;
    e.veh = v;
    e.next = next;
    e.rpos = rpos;
    e.logi = logi;
    if (q.size == 0) {
        q_insert(q, 0, e);
    } else {
        // "while (viter->first > ttime && viter != begin) --viter; insert(++viter)": behind the last entry that is not
        // later than the newcomer, but never in front of the head
        int i = q_find_last(q, e.exit, false);
        if (i < 0) i = 0;
        q_insert(q, i + 1, e);
    }
    q_close(q);
    mz_syncwarp();
    ++n_trav;
}

#ifdef MZ_WARP_COOP
// ---- register-resident queues (warp-per-node kernel) ----------------------------------------------------------------
// A queue of at most 32 entries is loaded ONCE, lane i holding list position i; every scan of Link::enter_veh and
// Q::exit_veh then is a ballot over registers and every list shift a store of the lane's own entry one slot further — no
// second trip to memory. (The generic code above re-reads the entries for each of its three scans and for the shift:
// ~6 dependent load rounds per successful turn, each an L2 round trip when the neighbour node has just written the queue.)
// Longer queues take the generic, chunked path.
constexpr int MZ_IMG = 32;
MZ_HD inline QEnt q_image(const QRef &q)
{
    QEnt e;
    if (mz_lane() < q.size) {
        e = ldx32(q_at(q, mz_lane()), q.sh);
    } else {
        e.exit = NAN;  // neither > t nor <= t
        e.entry = 0.0;
        e.veh = -1;
        e.next = (int)0x80000000;  // no link
        e.rpos = 0;
        e.logi = 0;
    }
    return e;
}
// Link::enter_veh on a queue whose image `img` is in registers (q.size <= MZ_IMG); `next` = the vehicle's next link
MZ_HD inline void link_enter_veh_img(const SimDev &d, QRef &q, const LinkStat &ls, const QEnt &img, int v, int rpos, int logi,
                                     int next, double t, unsigned long long &n_trav)
{
    const int lane = mz_lane();
    const bool valid = lane < q.size;
    // nr_running: elements from the first one with exit_time > t to the end (B/q.h:67-68)
    const unsigned run = mz_ballot(valid && img.exit > t);
    const int nr = run ? q.size - (mz_ffs(run) - 1) : 0;
    // calc_space: vehicle lengths from the back while exit_time <= t (B/q.cpp:346-361), summed back to front
    double space = 0.0;
    const int last_running = run ? 31 - mz_clz(run) : -1;
    if (last_running + 1 < q.size) {
        const double len = (valid && lane > last_running) ? lds(d.trip_length + img.veh) : 0.0;
        for (int k = q.size - 1; k > last_running; --k) space += mz_shfl(len, k);
    }
    const double queue_space = space / ls.lanes;
    const double running_length = ls.length - queue_space;
    double ro = 0.0;
    if (nr && running_length) ro = nr / (running_length * (ls.lanes / 1000.0));
    const double speed = sd_speed(lds(d.sd + link_sdfunc(q, ls)), ro);
    QEnt e;
    e.exit = t + (ls.length / speed);
    e.entry = t;
    e.veh = v;
    e.next = next;
    e.rpos = rpos;
    e.logi = logi;
    // Q::enter_veh (B/q.cpp:52-72): behind the last entry that is not later than the newcomer, never in front of the head
    int pos = 0;
    if (q.size > 0) {
        const unsigned le = mz_ballot(valid && img.exit <= e.exit);
        const int i = le ? 31 - mz_clz(le) : 0;
        pos = i + 1;
    }
    if (pos >= q.size - pos) {  // closer to the tail: the tail part moves back by one
        if (valid && lane >= pos) q_store(q, lane + 1, img);
        q.size++;
    } else {  // closer to the head: the head moves one slot forward, the front part follows it
        q.head = q.head == 0 ? q.cap - 1 : q.head - 1;
        q.size++;
        if (lane < pos) q_store(q, lane, img);
    }
    if (lane == 0) q_store(q, pos, e);
    q_close(q);
    mz_syncwarp();
    ++n_trav;
}
#endif

MZ_HD inline void link_enter_veh(const SimDev &d, QRef &q, const LinkStat &ls, int v, int rpos, int logi, double t,
                                 unsigned long long &n_trav)
{
#ifdef MZ_WARP_COOP
    if (q.size <= MZ_IMG) {
        mz_syncwarp();
        const int next = lds(d.route_term + rpos + 1);
        const QEnt img = q_image(q);
        link_enter_veh_img(d, q, ls, img, v, rpos, logi, next, t, n_trav);
        return;
    }
#endif
    link_enter_veh_generic(d, q, ls, v, rpos, logi, t, n_trav);
}

// bookkeeping of a successful Link::exit_veh (B/link.cpp:536-581 / 601-648) + exit log
MZ_HD inline void link_exit_bookkeeping(const SimDev &d, int l, const QEnt &e, double t, unsigned long long &n_exits)
{
    LinkAcc a = ldp32(d.lacc + l);
    const double traveltime = t - e.entry;
    a.avg_time = (a.nr_passed * a.avg_time + traveltime) / (a.nr_passed + 1);
    const int cp = a.curr_period;
    if ((cp + 1) * d.periodlength < e.entry) {
        double ta = a.tmp_avg;
        if (ta == 0.0) ta = cp < d.n_periods ? lds(d.link_hist + (size_t)l * d.n_periods + cp) : 0.0;
        if (cp < d.n_periods) wr(d.avgtimes + (size_t)l * d.n_periods + cp, ta);
        a.curr_period = cp + 1;
        a.tmp_avg = 0.0;
        a.tmp_passed = 0;
    } else {
        a.tmp_avg = (a.tmp_passed * a.tmp_avg + traveltime) / (a.tmp_passed + 1);
        a.tmp_passed += 1;
    }
    a.nr_passed += 1;
    if (mz_lane() == 0) {
        d.lacc[l] = a;
        d.exit_log[e.logi] = make_double2(e.entry, t);
    }
    ++n_exits;
}

// Link::update_exit_times + Q::update_exit_times (B/link.cpp:320-390, B/q.cpp:74-129); rare (only on unblocking).
// q = the in-link (opened), qn/lsn = the out-link that just unblocked.
MZ_HD MZ_COLD inline void link_update_exit_times(const SimDev &d, QRef &q, const LinkStat &ls, double t, int next,
                                         const QRef &qn, const LinkStat &lsn, int lookback)
{
    const double kjam = 142.0;
    const SdRec f = lds(d.sd + link_sdfunc(q, ls));
    const double k_dn = qn.size / (lsn.lanes * (lsn.length / 1000.0));  // Link::density
    const double v_exit = sd_speed(f, kjam);
    double v_dn = sd_speed(lds(d.sd + link_sdfunc(qn, lsn)), k_dn);
    double v_sw;
    if (v_dn <= v_exit) v_dn = v_exit - 1.0;
    if (k_dn >= f.romax) v_sw = v_exit - 1.0;
    else v_sw = (k_dn * v_dn) / (kjam - k_dn);
    if (v_sw >= v_exit) v_sw = v_exit - 1.0;
    if (v_sw == 0.0) v_sw = v_exit;
    // Q::update_exit_times walks the arrived prefix of the list with a running clock t0. The lanes load 32 entries
    // (exit, next, vehicle length) together; the walk itself — inherently sequential — then runs over registers.
    // `grp` counts the remaining members of a beyond-look-back lane group, which share one new exit time and are
    // overwritten without the "still running?" test (B/q.cpp:100-120).
    int cnt = 0, grp = 0;
    double t0 = t, grp_time = 0.0;
    bool stop = false;
    for (int base = 0; base < q.size && !stop; base += MZ_W) {
        const int i = base + mz_lane();
        double ex = 0.0, vl = 0.0;
        int nx = -1;
        QEnt *e = nullptr;
        if (i < q.size) {
            e = q_at(q, i);
            ex = ldx(&e->exit, q.sh);
            nx = ldx(&e->next, q.sh);
            vl = lds(d.trip_length + ldx(&e->veh, q.sh));
        }
        double mine = ex;  // this lane's entry: new exit time (written back if it changed)
        const int nchunk = q.size - base < MZ_W ? q.size - base : MZ_W;
        for (int k = 0; k < nchunk; ++k) {
            const double exk = mz_shfl(ex, k), vlk = mz_shfl(vl, k);
            const int nxk = mz_shfl(nx, k);
            double newk = exk;
            if (grp > 0) {
                newk = grp_time;
                --grp;
            } else {
                if (exk > t0) {
                    stop = true;
                    break;
                }
                if (cnt < lookback) {
                    cnt++;
                    if (nxk == next) {
                        newk = t0 + vlk / v_sw + vlk / v_exit;
                        t0 = newk;
                    }
                } else {
                    cnt++;
                    newk = t0 + vlk / v_sw + vlk / v_exit;
                    grp_time = newk;
                    grp = ls.lanes - 1;
                    t0 = newk;
                }
            }
            if (mz_lane() == k) mine = newk;
        }
        if (e && mine != ex) q_store_exit(q, i, mine);
        mz_syncwarp();
    }
    if (q.blocked_until == -2.0) {
        q.blocked_until = (ls.length / v_sw) + t;
        hot_store_blocked_until(q, q.blocked_until);
    }
    mz_syncwarp();
}

struct Tally {
    unsigned long long trav = 0, exits = 0, events = 0, arrived = 0;
};

// ---- transit overlay hooks (row f3) ---------------------------------------------------------------------------------
// Link::density_running_only (B/link.cpp:670-687) of an opened link, any queue length
MZ_HD MZ_COLD inline double link_density_running_only(const SimDev &d, const QRef &q, const LinkStat &ls, double t)
{
    mz_syncwarp();
    const int nr = q.size - q_find_first_running(q, t);
    double space = 0.0;
    if (q.size > 0) {
        const int last_running = q_find_last(q, t, true);
        for (int hi = q.size; hi > last_running + 1; hi -= MZ_W) {
            const int i = hi - 1 - mz_lane();
            const int veh = i > last_running ? ldx(&q_at(q, i)->veh, q.sh) : -1;
            const double len = veh >= 0 ? lds(d.trip_length + veh) : 0.0;
            const int left = hi - 1 - last_running;
            const int cnt = left < MZ_W ? left : MZ_W;
            for (int k = 0; k < cnt; ++k) space += mz_shfl(len, k);  // back to front, the reference's order (B/q.cpp:346-361)
        }
    }
    const double queue_space = space / ls.lanes;
    const double running_length = ls.length - queue_space;
    double ro = 0.0;
    if (nr && running_length) ro = nr / (running_length * (ls.lanes / 1000.0));
    return ro;
}
// is link l the link of bus v's next stop? (cars: one static load pair, no stops)
MZ_HD inline bool bus_stops_here(const SimDev &d, int v, int l)
{
    const int s0 = lds(d.trip_stop_off + v), s1 = lds(d.trip_stop_off + v + 1);
    if (s0 == s1) return false;
    const int k = ldm(d.bus_next + s0);
    return s0 + k < s1 && lds(d.stop_link + s0 + k) == l;
}
MZ_HD inline void bus_book_visit(const SimDev &d, int v, int k, int l, double t_enter, double t_arrive)
{
    if (mz_lane() == 0) {
        const unsigned long long i = MZ_ATOMIC_ADD(d.counters + 6, 1);
        StopVisit r;
        r.veh = v;
        r.stop = k;
        r.link = l;
        r.pad = 0;
        r.t_enter = t_enter;
        r.t_arrive = t_arrive;
        d.stop_out[i] = r;
    }
}
// Link::enter_veh for a bus whose next stop lies on this link (B/link.cpp:489-516): the exit time is computed like any
// vehicle's, the stop visit is booked at the stop's share of it, and the bus does NOT enter the queue
MZ_HD MZ_COLD inline void link_enter_bus(const SimDev &d, const QRef &q, const LinkStat &ls, int l, int v, int rpos, int logi, double t,
                                         unsigned long long &n_trav)
{
    const double ro = link_density_running_only(d, q, ls, t);
    const double speed = sd_speed(lds(d.sd + link_sdfunc(q, ls)), ro);
    const double exit_time = t + (ls.length / speed);
    const int s0 = lds(d.trip_stop_off + v);
    const int k = ldm(d.bus_next + s0);
    const double pos = lds(d.stop_pos + s0 + k);
    const double t_arrive = t + ((exit_time - t) * (pos / ls.length));
    if (mz_lane() == 0) {
        BusAt b;
        b.entry = t;
        b.rpos = rpos;
        b.logi = logi;
        d.bus_at[s0] = b;
        d.bus_next[s0] = k + 1;
    }
    bus_book_visit(d, v, k, l, t, t_arrive);
    mz_syncwarp();
    ++n_trav;
}
// Busstop::execute, "bus leaves the stop" (B/busline.cpp:1196-1263), on stop link l (sl = its index among the stop links):
// the link's density and speed NOW give the travel time of the whole link; the bus gets the share of the stretch in front
// of it — up to its next stop if that lies on this link too (another visit is booked), otherwise to the end of the link,
// where Q::enter_veh files it by that exit time. Returns the time of the link's next booked departure.
MZ_HD MZ_COLD inline double stop_execute(const SimDev &d, int l, int sl, double t)
{
    const int cur = ldp(d.slink_cur + sl);
    const int i = lds(d.dep_off + sl) + cur, i_end = lds(d.dep_off + sl + 1);
    const int v = lds(d.dep_veh + i);
    const LinkStat ls = lds(d.lstat + l);
    QRef q = q_open(d, l, ls);
    const double ro = link_density_running_only(d, q, ls, t);
    const double speed = sd_speed(lds(d.sd + link_sdfunc(q, ls)), ro);
    const double link_total_travel_time = ls.length / speed;
    const int s0 = lds(d.trip_stop_off + v), s1 = lds(d.trip_stop_off + v + 1);
    const int k = ldm(d.bus_next + s0);  // stops booked so far: the bus stands at stop k - 1
    const double position = lds(d.stop_pos + s0 + k - 1);
    BusAt b;
    b.entry = ldm(&d.bus_at[s0].entry);
    b.rpos = ldm(&d.bus_at[s0].rpos);
    b.logi = ldm(&d.bus_at[s0].logi);
    if (s0 + k < s1 && lds(d.stop_link + s0 + k) == l) {
        const double relative_length = (lds(d.stop_pos + s0 + k) - position) / ls.length;
        wr(d.bus_next + s0, k + 1);
        bus_book_visit(d, v, k, l, b.entry, t + link_total_travel_time * relative_length);
    } else if (q.size >= ls.maxcap) {
        if (mz_lane() == 0) MZ_ATOMIC_ADD(d.counters + 7, 1);  // Q::enter_veh returns false: the reference loses the bus
    } else {
        const double relative_length = (ls.length - position) / ls.length;
        QEnt e;
        e.exit = t + link_total_travel_time * relative_length;
        e.entry = b.entry;
        e.veh = v;
        e.next = lds(d.route_term + b.rpos + 1);
        e.rpos = b.rpos;
        e.logi = b.logi;
        if (q.size == 0) {
            q_insert(q, 0, e);
        } else {
            int p = q_find_last(q, e.exit, false);
            if (p < 0) p = 0;
            q_insert(q, p + 1, e);
        }
        q_close(q);
    }
    wr(d.slink_cur + sl, cur + 1);
    mz_syncwarp();
    return i + 1 < i_end ? lds(d.dep_t + i + 1) : INFINITY;
}

// Q::exit_veh(time, nextlink, lookback) on queues of any length — first vehicle in LIST order heading for `out`
// (B/q.cpp:156-222) — and, if it may leave, the hand-over to the out-link. Sets `ok` (a vehicle moved) or `nexttime`.
template <bool FEAT>
MZ_HD inline void turn_exit_generic(const SimDev &d, const TurnStat &ts, double t, Tally &ty, QRef &qi, QRef &qo, const LinkStat &lin,
                                    const LinkStat &lout, double &nexttime, bool &ok)
{
    const int n = q_find_first_next(qi, ts.out);
    if (n == qi.size) {
        nexttime = t + lin.freeflow;  // next_action = time + freeflowtime
    } else {
        const QEnt e = ldx32(q_at(qi, n), qi.sh);
        if (e.exit <= t) {
            if (n < ts.lookback) {
                q_erase(qi, n);
                q_close(qi);
                link_exit_bookkeeping(d, ts.in, e, t, ty.exits);
                if (FEAT && d.trip_stop_off && bus_stops_here(d, e.veh, ts.out))
                    link_enter_bus(d, qo, lout, ts.out, e.veh, e.rpos + 1, e.logi + 1, t + lds(&d.srv[ts.server].delay), ty.trav);
                else
                    link_enter_veh(d, qo, lout, e.veh, e.rpos + 1, e.logi + 1, t + lds(&d.srv[ts.server].delay), ty.trav);
                ok = true;
            } else {
                nexttime = t + 1.0 * (n / lout.lanes);  // integer division, B/q.cpp:204
            }
        } else {
            nexttime = e.exit;
        }
    }
}
#ifdef MZ_WARP_COOP
// the same for the warp-per-node kernel, as a real function: it re-opens the two queues (what turn_execute wrote so far is
// visible to the warp after the barrier) so that nothing of the caller's register state has to live in memory
template <bool FEAT>
MZ_HD MZ_COLD inline double turn_execute_long(const SimDev &d, const TurnStat &ts, double t, Tally &ty)
{
    mz_syncwarp();
    const LinkStat lin = lds(d.lstat + ts.in), lout = lds(d.lstat + ts.out);
    QRef qo = q_open(d, ts.out, lout);
    QRef qi = q_open(d, ts.in, lin);
    double nexttime = t;
    bool ok = false;
    turn_exit_generic<FEAT>(d, ts, t, ty, qi, qo, lin, lout, nexttime, ok);
    if (ok) {
        const SrvStat sv = lds(d.srv + ts.server);
        long long seed = ldp(d.srv_seed + ts.server);
        double mu, sd;
        server_rates_at(d, sv, t, mu, sd);
        const double nt = server_next(sv.type, d.server_type, mu, sd, sv.delay, seed, t);
        wr(d.srv_seed + ts.server, seed);
        return nt;
    }
    return nexttime < t ? t + 0.1 : nexttime;
}
#endif

// TurnAction::execute + Turning::process_veh (B/turning.cpp:237-265, 45-146); returns the re-booking time.
// FEAT = the kernels' SIG parameter: the network has signal controls or transit stops (their checks are compiled out otherwise)
template <bool FEAT>
MZ_HD inline double turn_execute(const SimDev &d, int k, const TurnStat &ts, double t, Tally &ty)
{
    const LinkStat lin = lds(d.lstat + ts.in), lout = lds(d.lstat + ts.out);
    QRef qo = q_open(d, ts.out, lout);
    QRef qi = q_open(d, ts.in, lin);
    const bool was_blocked = ldp(d.turn_blocked + k) != 0;
    double nexttime = t;
    bool ok = false;
    if (link_full_t(qo, lout, t)) {
        if (!was_blocked) {
            wr(d.turn_blocked + k, (unsigned char)1);
            hot_store_nr_exits_blocked(qi, qi.nr_exits_blocked + 1);
        }
        nexttime = t + 1.0;
    } else {
        if (was_blocked) {
            wr(d.turn_blocked + k, (unsigned char)0);
            link_update_exit_times(d, qi, lin, t, ts.out, qo, lout, ts.lookback);
            hot_store_nr_exits_blocked(qi, qi.nr_exits_blocked - 1);
        }
        if (qi.size == 0) {
            nexttime = t + lin.freeflow;
#ifdef MZ_WARP_COOP
        } else if (qi.size <= MZ_IMG && qo.size <= MZ_IMG) {
            // both queues in registers: one load round serves Q::exit_veh's search, the erase, and all of Link::enter_veh
            mz_syncwarp();
            const QEnt ii = q_image(qi), io = q_image(qo);
            const unsigned hit = mz_ballot(ii.next == ts.out);
            if (!hit) {
                nexttime = t + lin.freeflow;  // next_action = time + freeflowtime
            } else {
                const int n = mz_ffs(hit) - 1;
                QEnt e;
                e.exit = mz_shfl(ii.exit, n);
                if (e.exit <= t) {
                    if (n < ts.lookback) {
                        e.entry = mz_shfl(ii.entry, n);
                        e.veh = mz_shfl(ii.veh, n);
                        e.rpos = mz_shfl(ii.rpos, n);
                        e.logi = mz_shfl(ii.logi, n);
                        e.next = ts.out;
                        // loads of the rest of the event first, then the server draw (~1 k cycles of log / sqrt / sin) in their shadow
                        const int next2 = lds(d.route_term + e.rpos + 2);
                        const SrvStat sv = lds(d.srv + ts.server);
                        long long seed = ldp(d.srv_seed + ts.server);
                        double mu, sd;
                        server_rates_at(d, sv, t, mu, sd);
                        const double nt = server_next(sv.type, d.server_type, mu, sd, sv.delay, seed, t);
                        wr(d.srv_seed + ts.server, seed);
                        // erase list position n: the n entries in front of it move one slot towards the tail, the head advances
                        if (mz_lane() < n) q_store(qi, mz_lane() + 1, ii);
                        qi.head = qi.head + 1 == qi.cap ? 0 : qi.head + 1;
                        qi.size--;
                        q_close(qi);
                        link_exit_bookkeeping(d, ts.in, e, t, ty.exits);
                        if (FEAT && d.trip_stop_off && bus_stops_here(d, e.veh, ts.out))
                            link_enter_bus(d, qo, lout, ts.out, e.veh, e.rpos + 1, e.logi + 1, t + sv.delay, ty.trav);
                        else
                            link_enter_veh_img(d, qo, lout, io, e.veh, e.rpos + 1, e.logi + 1, next2, t + sv.delay, ty.trav);
                        return nt;
                    }
                    nexttime = t + 1.0 * (n / lout.lanes);  // integer division, B/q.cpp:204
                } else {
                    nexttime = e.exit;
                }
            }
#endif
        } else {
#ifdef MZ_WARP_COOP
            // a queue longer than the register image: rare (heavy congestion), kept out of the hot instruction stream
            return turn_execute_long<FEAT>(d, ts, t, ty);
#else
            turn_exit_generic<FEAT>(d, ts, t, ty, qi, qo, lin, lout, nexttime, ok);
#endif
        }
    }
    if (ok) {
        const SrvStat sv = lds(d.srv + ts.server);
        long long seed = ldp(d.srv_seed + ts.server);
        double mu, sd;
        server_rates_at(d, sv, t, mu, sd);
        const double nt = server_next(sv.type, d.server_type, mu, sd, sv.delay, seed, t);
        wr(d.srv_seed + ts.server, seed);
        return nt;
    }
    return nexttime < t ? t + 0.1 : nexttime;
}

// Daction::execute (B/node.cpp:637-673) with Q::exit_veh(time) (B/q.cpp:227-249)
MZ_HD MZ_COLD inline double dest_execute(const SimDev &d, int dst_server_slot, int svi, int l, double t, Tally &ty)
{
    const LinkStat ls = lds(d.lstat + l);
    QRef q = q_open(d, l, ls);
    if (q.size == 0) return t + ls.freeflow;
    const QEnt e = ldx32(q_at(q, 0), q.sh);
    if (e.exit <= t) {
        q_erase(q, 0);
        q_close(q);
        link_exit_bookkeeping(d, l, e, t, ty.exits);
        wr(d.arrival + e.veh, t);  // Vehicle::report(time)
        ++ty.arrived;
        long long seed = ldp(d.srv_seed + dst_server_slot);
        double nt;
        if (svi >= 0) {
            const SrvStat sv = lds(d.srv + svi);
            double mu, sd;
            server_rates_at(d, sv, t, mu, sd);
            nt = server_next(sv.type, d.server_type, mu, sd, sv.delay, seed, t);
        } else {
            nt = server_next(1, d.server_type, 0.1, 0.5, 0.001, seed, t);  // Destination's own Server(-20000,0,0.1,0.5,0.001)
        }
        wr(d.srv_seed + dst_server_slot, seed);
        return nt;
    }
    return e.exit;
}

// Origin::insert_veh / transfer_veh with the InputLink virtual queue (B/node.cpp:126-250, B/link.cpp:949-995).
// The virtual queue of origin o is the set of its trips in [vq_head, slot) whose `parked` flag is set, in trip order.
MZ_HD MZ_COLD inline void origin_insert(const SimDev &d, int o, int slot, double t, Tally &ty)
{
    const int *trips = d.origin_trips;
    const int v = lds(trips + slot);
    const int rpos0 = lds(d.trip_rpos0 + v);
    const int first = lds(d.route_term + rpos0);
    const LinkStat ls = lds(d.lstat + first);
    QRef q = q_open(d, first, ls);
    int vqh = ldp(d.origin_vq_head + o);
    while (vqh < slot && !ldp(d.trip_parked + lds(trips + vqh))) ++vqh;  // drop leading non-parked trips
    const bool vq_empty = vqh == slot;
    if (q.size >= ls.maxcap) {  // link->full(): park
        wr(d.trip_parked + v, (unsigned char)1);
    } else if (vq_empty) {
        if (d.trip_stop_off && bus_stops_here(d, v, first))  // a bus whose first stop lies on the first link of its route
            link_enter_bus(d, q, ls, first, v, rpos0, (int)lds(d.trip_log_off + v), t, ty.trav);
        else
            link_enter_veh(d, q, ls, v, rpos0, (int)lds(d.trip_log_off + v), t, ty.trav);
    } else {
        wr(d.trip_parked + v, (unsigned char)1);
        mz_syncwarp();
        // transfer_veh: head-first, every parked vehicle whose first link is `first`, while the link is not full
        for (int i = vqh; i <= slot; ++i) {
            if (q.size >= ls.maxcap) break;
            const int w = lds(trips + i);
            if (i != slot && !ldp(d.trip_parked + w)) continue;  // (the newcomer's flag was set just above)
            const int rp = lds(d.trip_rpos0 + w);
            if (lds(d.route_term + rp) != first) continue;
            wr(d.trip_parked + w, (unsigned char)0);
            if (d.trip_stop_off && bus_stops_here(d, w, first))
                link_enter_bus(d, q, ls, first, w, rp, (int)lds(d.trip_log_off + w), t, ty.trav);
            else
                link_enter_veh(d, q, ls, w, rp, (int)lds(d.trip_log_off + w), t, ty.trav);
        }
    }
    wr(d.origin_vq_head + o, vqh);
}

// Incident::execute without information broadcast (B/network.cpp:8092-8112) on link l: Link::set_incident at the start —
// the incident's sdfunc replaces the link's, the old one is kept in temp_sdfunc (its `blocked_until = -2.0` assigns the
// parameter, so the link's state is not touched) — and Link::unset_incident at the stop time (B/link.cpp:854-868).
MZ_HD MZ_COLD inline double incident_execute(const SimDev &d, int l, int inc, double t)
{
    const LinkStat ls = lds(d.lstat + l);
    QRef q = q_open(d, l, ls);
    const double stop = lds(d.inc_stop + inc);
    double nt;
    if (t < stop) {
        hot_store_sdfunc(q, lds(d.inc_sdfunc + inc) + 1, q.sd_cur ? q.sd_cur : ls.sdfunc + 1);
        nt = stop;
    } else {
        hot_store_sdfunc(q, q.sd_tmp, q.sd_tmp);
        hot_store_blocked_until(q, -1.0);
        nt = INFINITY;
    }
    mz_syncwarp();
    return nt;
}

// ------------------------------------------------------------------------------------------------ the window
// earliest pending action of node n (lanes scan the node's action list); `which` = -1 if the node has none
MZ_HD MZ_COLD inline Key node_min_key(const SimDev &d, int a0, int a1, int &which)
{
    Key best;
    best.t = INFINITY;
    best.tp = INFINITY;
    best.sub = 0x7fffffff;
    best.id = 0x7fffffff;
    for (int a = a0 + mz_lane(); a < a1; a += MZ_W * MZ_U) {
        ActDyn ad[MZ_U];
MZ_UNROLL
        for (int u = 0; u < MZ_U; ++u)
            if (a + u * MZ_W < a1) ad[u] = ldp32(d.adyn + a + u * MZ_W);
MZ_UNROLL
        for (int u = 0; u < MZ_U; ++u) {
            if (a + u * MZ_W >= a1) break;
            Key k;
            k.t = ad[u].t;
            k.tp = ad[u].tp;
            k.sub = ad[u].sub;
            k.id = a + u * MZ_W;
            if (key_less(k, best)) best = k;
        }
    }
    best = key_warp_min(best);
    which = best.id == 0x7fffffff ? -1 : best.id;
    return best;
}

// executes the earliest event of node n (identified by `which`); re-books that action and returns its new time.
// `sub` = the event's Lamport counter; a signal event advances it past the counters it hands to the executions it spawns.
// SIG = false compiles the signal handling out (the kernels for networks without signal controls are exactly the code
// they were before signals existed).
template <bool SIG>
MZ_HD inline double node_execute(const SimDev &d, int which, double t, int &sub, Tally &ty)
{
#if defined(MZ_SIM_PROFILE) && defined(__CUDACC__)
    const long long pc0 = clock64();
    const unsigned long long trav0 = ty.trav, arr0 = ty.arrived;
#endif
    ++ty.events;
    const ActStat as = lds(d.astat + which);
#ifdef MZ_EMU_TRACE
    mz_emu_trace(t, as.kind, as.idx);
#endif
    double nt;
    if (as.kind == ACT_TURN || (SIG && as.kind == ACT_SIGNAL)) {
        // A TurnAction event, or one signal event of control as.idx (ACT_SIGNAL): Stage::stop / Turning::activate for the
        // turnings the host listed (sim_host.h), after which the control's next event becomes this action's pending one.
        // Turning::activate (B/turning.cpp:218-227) executes every TurnAction of the turning on the spot, 3e-8 s apart, and
        // books it again — into a free event-chain slot of that action, keyed as a booking made by this event. Each
        // spawned execution takes the next Lamport counter of the instant, so the chains keep the reference's insertion
        // order among themselves and precede the control's own next event. Both cases share ONE inlined turn_execute.
        const bool sig = SIG && as.kind == ACT_SIGNAL;
        int k = as.idx, e0 = 0, e1 = 0, e = 0, o = 0, j = 0, na = 0, s0 = 0, cur = sub;
        TurnStat ts;
        ts.in = as.in;
        ts.out = as.out;
        ts.server = as.server;
        ts.lookback = as.lookback;
        SigEv ev;
        ev.n_ops = 0;
        ev.op_off = 0;
        double targ = t;
        if (sig) {
            e0 = lds(d.sigc_ev0 + k);
            e1 = lds(d.sigc_ev0 + k + 1);
            e = e0 + ldp(d.sig_next + k);
            ev = lds(d.sig_ev + e);
        }
        nt = INFINITY;
        for (;;) {
            if (sig) {
                while (j >= na && o < ev.n_ops) {  // next operation of the event
                    const int op = lds(d.sig_ops + ev.op_off + o);
                    ++o;
                    k = op >> 1;
                    ts = lds(d.turn + k);
                    wr(d.turn_active + k, (unsigned char)(op & 1));  // Turning::stop: red, pending chains die at their next event
                    j = 0;
                    na = 0;
                    if (op & 1) {
                        s0 = lds(d.turn_slot0 + k);
                        na = lds(d.turn_nact + k);
                        targ = t;
                    }
                }
                if (j >= na) break;  // every operation done
            } else if (SIG && d.turn_active && !ldp(d.turn_active + k)) {
                break;  // TurnAction::execute: a red turning's action does nothing and is not booked again (B/turning.cpp:236-262)
            }
            const double r = turn_execute<SIG>(d, k, ts, targ, ty);
            if (!sig) {
                nt = r;
                break;
            }
            int slot = -1;
            for (int q = 0; q < MZ_SIG_CHAINS && slot < 0; ++q)
                if (ldp(&d.adyn[s0 + j * MZ_SIG_CHAINS + q].t) == INFINITY) slot = q;
            if (slot < 0) {  // more live chains than slots: cannot be represented — flag it, the host reports the failure
                if (mz_lane() == 0) d.counters[5] = 3;
                slot = 0;
            }
            ++cur;
            if (mz_lane() == 0) {
                ActDyn ad;
                ad.t = r;
                ad.tp = t;
                ad.sub = cur;
                ad.pad = 0;
                ad.pad2 = 0.0;
                d.adyn[s0 + j * MZ_SIG_CHAINS + slot] = ad;
            }
            mz_syncwarp();
            ++j;
            targ += 0.00000003;
        }
        if (sig) {
            ++cur;
            wr(d.sig_endsub + e, cur);
            wr(d.sig_next + as.idx, e + 1 - e0);
            ActDyn own;
            own.t = INFINITY;
            own.tp = INFINITY;
            own.sub = 0;
            own.pad = 0;
            own.pad2 = 0.0;
            if (e + 1 < e1) {
                const SigEv nx = lds(d.sig_ev + e + 1);
                own.t = nx.t;
                own.tp = nx.tp;
                own.sub = nx.parent < 0 ? nx.sub0 : (nx.parent == e ? cur : ldp(d.sig_endsub + nx.parent));
            }
            if (mz_lane() == 0) d.adyn[which] = own;
            mz_syncwarp();
            sub = cur;
            return own.t;
        }
    } else if (as.kind == ACT_DEST) {
        nt = dest_execute(d, as.aux, as.aux2, as.idx, t, ty);
    } else if (as.kind == ACT_INCIDENT) {
        nt = incident_execute(d, as.idx, as.aux, t);
    } else if (SIG && as.kind == ACT_STOP) {
        nt = stop_execute(d, as.idx, as.aux, t);
    } else {  // ODaction::execute (B/od.cpp:69-108): the pair's next pre-generated trip enters at its origin
        const int k = as.idx;
        const int done = ldp(d.od_next + k);
        const int pos = lds(d.od_off + k) + done;
        const int v = lds(d.od_trips + pos);
        origin_insert(d, as.aux, lds(d.origin_trip_off + as.aux) + lds(d.trip_slot + v), t, ty);
        wr(d.od_next + k, done + 1);
        nt = pos + 1 < lds(d.od_off + k + 1) ? lds(d.trip_time + lds(d.od_trips + pos + 1)) : INFINITY;
    }
    if (mz_lane() == 0) {
        ActDyn ad;
        ad.t = nt;
        ad.tp = t;
        ad.sub = sub;
        ad.pad = 0;
        ad.pad2 = 0.0;
        d.adyn[which] = ad;
    }
    mz_syncwarp();
#if defined(MZ_SIM_PROFILE) && defined(__CUDACC__)
    if (mz_lane() == 0) {  // classes: 0 turn fail, 1 turn success, 2 sink fail, 3 sink success, 4 origin
        const int cls = as.kind == ACT_TURN ? (ty.trav > trav0 ? 1 : 0) : as.kind == ACT_DEST ? (ty.arrived > arr0 ? 3 : 2) : 4;
        const unsigned long long dt = (unsigned long long)(clock64() - pc0);
        atomicAdd(d.counters + 8 + 3 * cls, dt);
        atomicAdd(d.counters + 9 + 3 * cls, 1ull);
        atomicMax(d.counters + 10 + 3 * cls, dt);
    }
#endif
    return nt;
}

// warp-wide minimum of a non-negative double
MZ_HD inline double warp_min_time(double t)
{
#ifdef MZ_WARP_COOP
    const unsigned long long tb = (unsigned long long)__double_as_longlong(t);
    const unsigned hi = (unsigned)(tb >> 32), lo = (unsigned)tb;
    const unsigned mh = __reduce_min_sync(0xffffffffu, hi);
    const unsigned ml = __reduce_min_sync(0xffffffffu, hi == mh ? lo : 0xffffffffu);
    return __longlong_as_double((long long)(((unsigned long long)mh << 32) | ml));
#else
    for (int m = MZ_W / 2; m > 0; m >>= 1) {
        const double o = mz_shfl_xor(t, m);
        t = o < t ? o : t;
    }
    return t;
#endif
}

enum { VISIT_WAIT = 0, VISIT_RAN = 1, VISIT_DONE = 2 };

// One visit of node n by one warp (host emulation: one lane): if the node's earliest event precedes every neighbour's
// published key, execute its events up to that bound and publish the new key. `final_visit` selects the clean-up pass
// that executes exactly one event whatever its time (the first event at/after the horizon).
// Returns VISIT_DONE when the node has nothing left inside the horizon; *key_time receives the node's (new) key time.
// `my_t` / `nb_bound` (>= 0): the caller's cached own key time and the smallest neighbour time it has just read — the
// visit then needs no key loads of its own; pass -1 to have them read here.
// `acc` (5 counters in the order of d.counters[0..4]): where the visit's tallies go; null = atomics on d.counters.
template <bool SIG = true>
MZ_HD MZ_NOINLINE int async_visit(const SimDev &d, int n, bool final_visit, double my_t, double nb_bound, double *key_time,
                                  unsigned long long *acc = nullptr)
{
#if defined(MZ_SIM_PROFILE) && defined(__CUDACC__)
    const long long vc0 = clock64();
#endif
    // the node's sharing class decides how its published state is written: block-, gpu- or system-scope release
    const int mr = node_class(d, n);
    Key mine;
    mine.t = my_t >= 0.0 ? my_t : node_key_time(d, n, mr);  // own key: only this warp ever writes it
    mine.id = n;
    *key_time = mine.t;
    if (!final_visit && !(mine.t < d.runtime)) return VISIT_DONE;
    const int nb0 = lds(d.node_nb_off + n), nb1 = lds(d.node_nb_off + n + 1);
    // the bound: the smallest published neighbour key. Times alone decide unless a neighbour ties with this node.
    double bt = nb_bound;
    if (!(nb_bound >= 0.0)) {
        bt = INFINITY;
        for (int i = nb0 + mz_lane(); i < nb1; i += MZ_W) {
            const int nbx = lds(d.node_nb + i);
            const double mt = node_key_time(d, nbx & MZ_NB_MASK, nbx >> MZ_NB_SHIFT);
            bt = mt < bt ? mt : bt;
        }
        bt = warp_min_time(bt);
    }
    Key bound;
    bound.t = bt;
    bound.tp = -INFINITY;  // "before every key of that instant": an event at exactly bt waits for the full keys
    bound.sub = (int)0x80000000;
    bound.id = (int)0x80000000;
    if (!final_visit) {
        if (bt < mine.t) return VISIT_WAIT;
        if (bt == mine.t) {
            const Key own = node_key_full(d, n, mr);
            mine.tp = own.tp;
            mine.sub = own.sub;
            bound.t = INFINITY;
            bound.tp = INFINITY;
            bound.sub = 0x7fffffff;
            bound.id = 0x7fffffff;
            for (int i = nb0 + mz_lane(); i < nb1; i += MZ_W) {
                const int nbx = lds(d.node_nb + i);
                const Key k = node_key_full(d, nbx & MZ_NB_MASK, nbx >> MZ_NB_SHIFT);
                if (key_less(k, bound)) bound = k;
            }
            bound = key_warp_min(bound);
            if (!key_less(mine, bound)) return VISIT_WAIT;
        }
    }
    // No neighbour can be executing now (it would need a key below this node's), and what each of them wrote before
    // publishing the key just read is visible: it is read past L1 or lives in this SM's L1 (see "memory access" above).
    mz_acquire(mr);
    // latest time instant at which a neighbour executed and the largest Lamport sub-counter used there
    const NodeLast own_last = node_last(d, n, mr);  // issued together with the neighbours' records
    double nb_t = 0.0;  // 0 = never: all event times are >= 0.1
    int nb_sub = 0;
    for (int i = nb0 + mz_lane(); i < nb1; i += MZ_W * MZ_U) {
        NodeLast l[MZ_U];
MZ_UNROLL
        for (int u = 0; u < MZ_U; ++u) {
            l[u].t = 0.0;
            l[u].sub = 0;
            if (i + u * MZ_W < nb1) {
                const int nbx = lds(d.node_nb + i + u * MZ_W);
                l[u] = node_last(d, nbx & MZ_NB_MASK, nbx >> MZ_NB_SHIFT);
            }
        }
MZ_UNROLL
        for (int u = 0; u < MZ_U; ++u)
            if (l[u].t > nb_t || (l[u].t == nb_t && l[u].sub > nb_sub)) {
                nb_t = l[u].t;
                nb_sub = l[u].sub;
            }
    }
    warp_max_instant(nb_t, nb_sub);
    const int a0 = lds(d.node_act_off + n), a1 = lds(d.node_act_off + n + 1);
    Tally ty;
    double last_t = own_last.t;
    int last_sub = own_last.sub;
    Key k;
    int which;
    // The node's action keys stay in registers across the events of this visit (lane j holds action a0+j): only the
    // executed action changes, so its lane updates its copy instead of re-reading all records after every event.
    // (a signal event rewrites the keys of OTHER actions of its node — their chain slots — so such nodes re-read them)
    const bool in_regs = a1 - a0 <= MZ_W && !(SIG && d.node_sig && lds(d.node_sig + n));
    Key mykey;
    mykey.t = INFINITY;
    mykey.tp = INFINITY;
    mykey.sub = 0x7fffffff;
    mykey.id = 0x7fffffff;
    if (in_regs && a0 + mz_lane() < a1) {
        mz_prefetch(d.astat + a0 + mz_lane());  // the static record of whichever action wins arrives with the keys
        const ActDyn ad = ldp32(d.adyn + a0 + mz_lane());
        mykey.t = ad.t;
        mykey.tp = ad.tp;
        mykey.sub = ad.sub;
        mykey.id = a0 + mz_lane();
    }
#if defined(MZ_SIM_PROFILE) && defined(__CUDACC__)
    const long long vc1 = clock64();
#endif
    for (int it = 0;; ++it) {
        if (in_regs) {
            k = key_warp_min(mykey);
            which = k.id == 0x7fffffff ? -1 : k.id;
        } else {
            k = node_min_key(d, a0, a1, which);
        }
        if (which < 0) break;
        if (final_visit) {
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
        const double nt = node_execute<SIG>(d, which, k.t, sub, ty);  // (a signal event advances sub past its spawned executions)
        last_sub = sub;
        if (in_regs && which == a0 + mz_lane()) {
            mykey.t = nt;
            mykey.tp = k.t;
            mykey.sub = sub;
        }
    }
#if defined(MZ_SIM_PROFILE) && defined(__CUDACC__)
    const long long vc2 = clock64();
#endif
    if (ty.events) {
        node_last_store(d, n, last_t, last_sub, own_last.ver + 1u);
        node_key_publish(d, n, k, own_last.ver, mr);
#if defined(MZ_SIM_PROFILE) && defined(__CUDACC__)
        if (mz_lane() == 0) {  // per sharing class of the node: visits, cycles before / inside / after the event loop
            const long long vc3 = clock64();
            atomicAdd(d.counters + 32 + 4 * mr, 1ull);
            atomicAdd(d.counters + 33 + 4 * mr, (unsigned long long)(vc1 - vc0));
            atomicAdd(d.counters + 34 + 4 * mr, (unsigned long long)(vc2 - vc1));
            atomicAdd(d.counters + 35 + 4 * mr, (unsigned long long)(vc3 - vc2));
        }
#endif
        if (acc) {  // the caller's private tallies (registers of a thread, or a warp's slot in shared memory)
            if (mz_lane() == 0) {
                acc[0] += ty.trav;
                acc[1] += ty.exits;
                acc[2] += ty.events;
                acc[3] += ty.arrived;
                acc[4] += 1;
            }
        } else if (mz_lane() == 0) {
            MZ_ATOMIC_ADD(d.counters + 2, ty.events);
            MZ_ATOMIC_ADD(d.counters + 4, 1);
            if (ty.trav) MZ_ATOMIC_ADD(d.counters + 0, ty.trav);
            if (ty.exits) MZ_ATOMIC_ADD(d.counters + 1, ty.exits);
            if (ty.arrived) MZ_ATOMIC_ADD(d.counters + 3, ty.arrived);
        }
    }
    *key_time = k.t;
    if (!ty.events) return VISIT_WAIT;
    return (k.t < d.runtime) ? VISIT_RAN : VISIT_DONE;
}

}  // namespace MZ_SIM_NS
