// Hot path 1: the link update on sm_100a — CUDA backend + C ABI (mz_sim_*).
// The data layout and per-event functions live in sim_core.h, the backend-independent host logic in sim_host.h.
#include "common.h"
#include "sim_host.h"

using namespace mzsim;

namespace mzgroup {  // sim_group.cu: the thread-per-node variant of the node-visit kernel (up to 32 nodes per warp at a time)
int launch(const void *simdev, size_t simdev_bytes, int max_region, unsigned long long max_idle, cudaStream_t stream);
int max_blocks(int device, int *blocks_out);
int threads_per_block();
}

namespace {

// dst[i * stride] = pattern for i < n (64-bit words): the parts of Network::reset that are not a byte pattern
__global__ void k_fill64(unsigned long long *dst, unsigned long long pattern, size_t n, size_t stride)
{
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) dst[i * stride] = pattern;
}

// Clean-up pass: exactly one event of one node (the first event at/after the horizon). One warp.
__global__ void __launch_bounds__(32) k_sim_final(const __grid_constant__ SimDev d, int node)
{
    double t;
    async_visit(d, node, true, -1.0, -1.0, &t);
}

// ---- the persistent node-visit loop ---------------------------------------------------------------------------------
// The whole run executes inside ONE kernel without any grid-wide barrier. A thread block owns one REGION of this rank's
// nodes (a contiguous patch of the road graph, sim_host.h: sim_set_regions): most of a node's neighbours and links belong
// to the same SM, so their keys, link records and queues are read through that SM's L1 and ordered with block-scope
// fences; only state on a region boundary goes past L1 to L2, and nothing ever invalidates L1 (sim_core.h, "memory access").
//
// Scheduling is EVENT-DRIVEN, not a sweep. The run time of the link update is its critical path — tens of thousands of
// dependent rounds "node runs, neighbour sees the new key, neighbour runs" — so what counts is how fast a node that has
// just become runnable is noticed (measured with the earlier sweep loop: at 1.02 M links a warp needed 42 us to come
// round to the same node again, four times the visit itself; only 2 % of the polled nodes were runnable). Each block keeps
// two bit masks over its region in shared memory: READY (somebody changed something this node may be waiting for) and
// BUSY (a warp is working on it). Any warp of the block takes any ready node: claim (BUSY), clear READY, read the
// neighbours' published key times — lanes = neighbours — and, if the node's key precedes them all, visit it
// (async_visit). A visit that published a new key then wakes exactly the neighbours it may have unblocked — those whose
// key time lies at or below the new key; adjacent nodes never run at the same time, so the key times read before the visit
// are still exact — by setting their READY bits: a shared-memory atomic inside the region, an L2 atomic into the other
// block's mailbox word across regions (after a second release fence, so that the key is visible before the wake-up).
// Nodes with a neighbour on another GPU cannot be woken by it and are re-marked ready periodically (they are polled, as
// every node was before). READY is cleared before the keys are read and set after a key is published, so no wake-up is
// lost; a slow backstop re-marks the region-boundary nodes all the same.
#ifndef MZ_PTHREADS
#define MZ_PTHREADS 384
#endif
constexpr int kPThreads = MZ_PTHREADS;
#ifndef MZ_THREAD_PER_NODE_FROM
#define MZ_THREAD_PER_NODE_FROM 600000  // measured (round 2): 270 k nodes 3091 ms (warp) vs 3247 ms (thread)
#endif
constexpr int kThreadPerNodeFrom = MZ_THREAD_PER_NODE_FROM;
struct alignas(16) NodeSlot {  // per node of the region: cached key time, node id (-1 = none), done flag
    double keyt;
    int node;
    int done;
};
constexpr int kSlotBytes = (int)sizeof(NodeSlot);
constexpr int kMaxRegion = ((200 * 1024) / (kSlotBytes * 32 + 8)) * 32;  // nodes a block can track (slots + two mask words per 32)

__device__ __forceinline__ unsigned lds_volatile(const unsigned *p) { return *reinterpret_cast<const volatile unsigned *>(p); }

// SIG: the network has signal controls or transit stops (sim_core.h compiles their handling out otherwise)
template <bool SIG>
__global__ void __launch_bounds__(kPThreads, 1) k_sim_notify(const __grid_constant__ SimDev d, unsigned long long max_idle_cycles, int cap)
{
    // dynamic shared memory: `cap` node slots (cap = the largest region rounded up to 32), then cap/32 READY and BUSY words
    extern __shared__ NodeSlot s_slots[];
    unsigned *s_ready = reinterpret_cast<unsigned *>(s_slots + cap);
    unsigned *s_busy = s_ready + cap / 32;
    __shared__ unsigned long long s_acc[kPThreads / 32][8];  // per warp: traversals, exits, events, arrived, visits
    __shared__ unsigned s_done, s_abort, s_progress;  // s_progress: visits of the block that published a key
    const unsigned lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    constexpr unsigned G = kPThreads / 32;
    if (lane < 8) s_acc[warp][lane] = 0;
    const unsigned r0 = (unsigned)lds(d.region_off + blockIdx.x), r1 = (unsigned)lds(d.region_off + blockIdx.x + 1);
    const unsigned n_reg = r1 - r0, nw = (n_reg + 31) / 32;
    if (n_reg > (unsigned)cap) {  // the host sizes the slots so that this cannot happen
        if (threadIdx.x == 0) d.counters[5] = 2;
        return;
    }
    if (threadIdx.x == 0) {
        s_done = 0;
        s_abort = 0;
        s_progress = 0;
    }
    for (unsigned p = threadIdx.x; p < (unsigned)cap; p += kPThreads) {
        NodeSlot ns;
        ns.node = p < n_reg ? lds(d.my_nodes + r0 + p) : -1;
        ns.keyt = ns.node >= 0 ? node_key_time(d, ns.node, 2) : INFINITY;
        ns.done = ns.node < 0;
        s_slots[p] = ns;
    }
    for (unsigned w = threadIdx.x; w < (unsigned)cap / 32; w += kPThreads) {
        const unsigned left = w * 32 < n_reg ? n_reg - w * 32 : 0;
        s_ready[w] = left >= 32 ? 0xffffffffu : ((1u << left) - 1u);  // every node starts ready
        s_busy[w] = 0;
    }
    __syncthreads();
    unsigned *mbox = d.mailbox + lds(d.mb_off + blockIdx.x);
    const int poll0 = lds(d.poll_off + blockIdx.x), poll1 = lds(d.poll_off + blockIdx.x + 1);
    const unsigned start = nw ? (warp * ((nw + G - 1) / G)) % nw : 0;  // the warps start their scans at different words
    long long last_progress = clock64();
    unsigned seen_progress = 0;
    unsigned iter = warp;
#ifdef MZ_SIM_PROFILE
    const long long kc0 = clock64();
    long long in_visit = 0, n_sweeps = 0, n_calls = 0, n_idle = 0;
#endif
    while (lds_volatile(&s_done) < n_reg && !lds_volatile(&s_abort)) {
        ++iter;
#ifdef MZ_SIM_PROFILE
        ++n_sweeps;
#endif
        // nodes on the cut between GPUs: nobody can wake them (their neighbour's block lives on another GPU)
        if (poll1 > poll0 && (iter & 3u) == 0)
            for (int i = poll0 + (int)lane; i < poll1; i += 32) {
                const int p = lds(d.poll_pos + i);
                atomicOr(&s_ready[p >> 5], 1u << (p & 31));
            }
        // ---- take a ready node that nobody is working on
        int pick = -1;
        unsigned pw = 0, pbit = 0;
        for (unsigned base = 0; base < nw; base += 32) {
            unsigned w = base + lane, c = 0;
            if (w < nw) {
                w += start;
                if (w >= nw) w -= nw;
                c = lds_volatile(&s_ready[w]) & ~lds_volatile(&s_busy[w]);
            }
            const unsigned m = __ballot_sync(0xffffffffu, c != 0);
            if (m) {
                const int src = __ffs((int)m) - 1;
                pw = __shfl_sync(0xffffffffu, w, src);
                pbit = 1u << (__ffs((int)__shfl_sync(0xffffffffu, c, src)) - 1);
                unsigned old = 0;
                if (lane == 0) {
                    old = atomicOr(&s_busy[pw], pbit);
                    if (!(old & pbit)) atomicAnd(&s_ready[pw], ~pbit);  // READY is cleared BEFORE the keys are read
                }
                old = __shfl_sync(0xffffffffu, old, 0);
                __syncwarp();
                if (!(old & pbit)) pick = (int)(pw * 32) + __ffs((int)pbit) - 1;
                break;  // (lost the claim to another warp: scan again)
            }
        }
        if (pick >= 0) {
            __threadfence_block();
            const NodeSlot ns = s_slots[pick];
            const double kt = ns.keyt;
            bool done_now = false, woke = false;
            if (!ns.done) {
                const int n = ns.node;
                if (!(kt < d.runtime)) {
                    done_now = true;  // nothing left inside the horizon
                } else {
                    // the neighbours' published key times, lane = neighbour
                    const int nb0 = lds(d.node_nb_off + n), nb1 = lds(d.node_nb_off + n + 1);
                    double nbt = INFINITY, mine_t = INFINITY;  // mine_t: key time of this lane's (first) neighbour
                    for (int i = nb0 + (int)lane; i < nb1; i += 32) {
                        const int nbx = lds(d.node_nb + i);
                        const double mt = node_key_time(d, nbx & MZ_NB_MASK, nbx >> MZ_NB_SHIFT);
                        if (i < nb0 + 32) mine_t = mt;
                        nbt = mt < nbt ? mt : nbt;
                    }
                    nbt = warp_min_time(nbt);
                    if (!(nbt < kt)) {
                        double nt;
#ifdef MZ_SIM_PROFILE
                        const long long vc = clock64();
#endif
                        const int st = async_visit<SIG>(d, n, false, kt, nbt, &nt, s_acc[warp]);
#ifdef MZ_SIM_PROFILE
                        in_visit += clock64() - vc;
                        ++n_calls;
#endif
                        if (st != VISIT_WAIT) {
                            if (lane == 0) s_slots[pick].keyt = nt;
                            done_now = st == VISIT_DONE;
                            woke = true;
                            // the key is published (lane 0 stored it): order it before the wake-ups of every lane
                            __threadfence_block();
                            __syncwarp();
                            bool remote = false;
                            for (int i = nb0 + (int)lane; i < nb1; i += 32) {
                                // only neighbours the new key has passed can have been waiting for this node
                                if (i < nb0 + 32 && mine_t > nt) continue;
                                const int sl = lds(d.nb_slot + i);
                                if (sl >= 0) atomicOr(&s_ready[sl >> 5], 1u << (sl & 31));
                                else if (sl <= -2) remote = true;
                            }
                            if (__any_sync(0xffffffffu, remote)) {
                                if (remote) {
                                    asm volatile("fence.release.gpu;" ::: "memory");  // key before wake-up, across SMs
                                    for (int i = nb0 + (int)lane; i < nb1; i += 32) {
                                        if (i < nb0 + 32 && mine_t > nt) continue;
                                        const int sl = lds(d.nb_slot + i);
                                        if (sl <= -2) {
                                            const unsigned g = (unsigned)(-(sl + 2));
                                            atomicOr(d.mailbox + (g >> 5), 1u << (g & 31));
                                        }
                                    }
                                }
                                __syncwarp();
                            }
                            // stopped although its next event is not later than the bound (a tie that needs the full
                            // keys, or the event cap): nobody else will wake it
                            if (st == VISIT_RAN && !(nbt < nt) && lane == 0) atomicOr(&s_ready[pw], pbit);
                        }
                    }
                }
            }
            if (lane == 0) {
                if (done_now) {
                    s_slots[pick].done = 1;
                    atomicAdd(&s_done, 1u);
                }
                if (woke || done_now) atomicAdd(&s_progress, 1u);
                __threadfence_block();
                atomicAnd(&s_busy[pw], ~pbit);
            }
            __syncwarp();
            continue;
        }
        // ---- nothing ready: collect the wake-ups other blocks left in this region's mailbox
#ifdef MZ_SIM_PROFILE
        ++n_idle;
#endif
        bool got = false;
        for (unsigned w = lane; w < nw; w += 32) {
            unsigned v;
            asm volatile("ld.relaxed.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(mbox + w) : "memory");
            if (v) {
                v = atomicExch(mbox + w, 0u);
                if (v) {
                    atomicOr(&s_ready[w], v);
                    got = true;
                }
            }
        }
        if (__any_sync(0xffffffffu, got)) {
            __threadfence_block();
            continue;
        }
        // backstop: once in a while re-mark every node that has a neighbour outside the region (costs a key check each)
        if ((iter & 0xfffu) == 0)
            for (unsigned p = lane; p < n_reg; p += 32) {
                const NodeSlot ns = s_slots[p];
                if (!ns.done && node_class(d, ns.node) != 0) atomicOr(&s_ready[p >> 5], 1u << (p & 31));
            }
        const unsigned prog = lds_volatile(&s_progress);
        if (prog != seen_progress) {
            seen_progress = prog;
            last_progress = clock64();
        }
        if ((unsigned long long)(clock64() - last_progress) > max_idle_cycles) {  // cannot happen: the node with the globally smallest key can always run
            if (lane == 0) {
                d.counters[5] = 1;
                s_abort = 1;
            }
            break;
        }
        if ((iter & 63u) == 0) {  // another block gave up: do not wait for its nodes for ever
            unsigned long long e;
            asm volatile("ld.relaxed.gpu.global.u64 %0, [%1];" : "=l"(e) : "l"(d.counters + 5) : "memory");
            if (e) break;
        }
        __nanosleep(40);
    }
    __syncwarp();
    if (lane < 5 && s_acc[warp][lane]) atomicAdd(d.counters + lane, s_acc[warp][lane]);
#ifdef MZ_SIM_PROFILE
    if (lane == 0) {  // warp totals: cycles alive, cycles inside async_visit, loop iterations, visit calls, warps
        atomicAdd(d.counters + 48, (unsigned long long)(clock64() - kc0));
        atomicAdd(d.counters + 49, (unsigned long long)in_visit);
        atomicAdd(d.counters + 50, (unsigned long long)n_sweeps);
        atomicAdd(d.counters + 51, (unsigned long long)n_calls);
        atomicAdd(d.counters + 52, 1ull);
        atomicAdd(d.counters + 53, (unsigned long long)n_idle);
    }
#endif
}

// ---- the sweep scheduler (MZ_SIM_OPT_KERNEL = 1, the default) -------------------------------------------------------
// A warp owns a fixed, strided subset of the region and keeps sweeping it: lane j pre-checks node j (cached own key time
// against the neighbours' published times, eight loads in flight), and every node that may be a local minimum gets a
// warp-cooperative visit (async_visit). Measured against the event-driven scheduler above on B200: config 2 212 vs 222 ms,
// 1.02 M links 3.09 vs 3.80 s per simulated hour — the pre-check of 32 nodes at once by the lanes of a warp costs less
// than one claim + key check per wake-up (3.8 of them per visit at 1.02 M links, each a chain of L2 / DRAM round trips).
constexpr int kSweepSlotBytes = 12;  // key time + node id
constexpr int kMaxGroups = (227 * 1024) / (kPThreads * kSweepSlotBytes) & ~1;  // x32 nodes a warp can own (dynamic shared memory)

// SIG: the network has signal controls or transit stops (sim_core.h compiles their handling out otherwise)
template <bool SIG>
__global__ void __launch_bounds__(kPThreads, 1) k_sim_async(const __grid_constant__ SimDev d, unsigned long long max_idle_sweeps, int groups)
{
    // per-warp state of the owned nodes (slot = group*32 + lane): node id, cached key time, done flag. Shared memory, so
    // the group loop stays a run-time loop and async_visit is inlined exactly once — with the kernel parameter `d`
    // addressed as constant-bank operands instead of through a local-memory copy (a noinline call cost ~4 k cycles per visit).
    // Dynamic shared memory: `groups` x 32 slots per warp.
    // (two arrays, 12 bytes per node: at 1.02 M links the slot table then stays below 32 KB and the SM keeps the next larger
    // L1 configuration — L1 capacity is worth more to this kernel than anything else shared memory could hold)
    extern __shared__ double s_dyn[];
    __shared__ unsigned long long s_acc[kPThreads / 32][8];  // per warp: traversals, exits, events, arrived, visits
    const unsigned lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (lane < 8) s_acc[warp][lane] = 0;
    constexpr unsigned G = kPThreads / 32;
    const unsigned S = (unsigned)groups * 32;
    // region of this block; its nodes idx = r0 + warp + j*G; a warp handles them in groups of 32 (lane = node of the group)
    const unsigned r0 = (unsigned)lds(d.region_off + blockIdx.x), r1 = (unsigned)lds(d.region_off + blockIdx.x + 1);
    const unsigned gw = r0 + warp;
    const unsigned n_own = gw < r1 ? (r1 - gw + G - 1) / G : 0;
    if (n_own > S) {  // the host sizes the slots so that this cannot happen
        if (lane == 0) d.counters[5] = 2;
        return;
    }
    double *s_keyt = s_dyn + warp * S;                                                        // cached key time
    int *s_node = reinterpret_cast<int *>(s_dyn + (size_t)(kPThreads / 32) * S) + warp * S;  // node id, -1 = none / retired
    for (unsigned j = lane; j < S; j += 32) {
        const int node = j < n_own ? lds(d.my_nodes + gw + j * G) : -1;
        s_node[j] = node;
        s_keyt[j] = node >= 0 ? node_key_time(d, node, 2) : INFINITY;
    }
    __syncwarp();
    const unsigned n_groups = (n_own + 31) / 32;
    unsigned n_done = 0;
    unsigned long long idle = 0;
#ifdef MZ_SIM_PROFILE
    const long long kc0 = clock64();
    long long in_visit = 0, n_sweeps = 0, n_calls = 0;
#endif
    while (n_done < n_own) {
#ifdef MZ_SIM_PROFILE
        ++n_sweeps;
#endif
        bool ran = false;
#pragma unroll 1
        for (unsigned g = 0; g < n_groups; ++g) {
            const unsigned slot = g * 32 + lane;
            bool cand = false;
            double nbt = INFINITY;  // smallest neighbour key time seen by this lane's pre-check
            const int my_node = s_node[slot];
            const double kt = s_keyt[slot];
            if (my_node >= 0) {
                cand = true;  // (a node past the horizon is a candidate too: the visit retires it)
                if (kt < d.runtime) {
                    const int n = my_node;
                    const int nb0 = lds(d.node_nb_off + n), nb1 = lds(d.node_nb_off + n + 1);
                    for (int i0 = nb0; i0 < nb1; i0 += 8) {
                        double mt[8];
#pragma unroll
                        for (int k = 0; k < 8; ++k) {
                            mt[k] = INFINITY;
                            if (i0 + k < nb1) {
                                const int nbx = lds(d.node_nb + i0 + k);
                                mt[k] = node_key_time(d, nbx & MZ_NB_MASK, nbx >> MZ_NB_SHIFT);
                            }
                        }
#pragma unroll
                        for (int k = 0; k < 8; ++k) nbt = mt[k] < nbt ? mt[k] : nbt;
                    }
                    cand = !(nbt < kt);
                    if (cand) {
                        // The warp visits its candidates one after the other: start fetching every candidate's action keys
                        // now, so that the later ones find them in L1 (on large networks they come from DRAM).
                        const int a0 = lds(d.node_act_off + n), a1 = lds(d.node_act_off + n + 1);
                        for (int p = a0; p < a1; p += 4) mz_prefetch(d.adyn + p);
                        // (Tried: remembering per slot the neighbour that blocked the node last time and checking that one first
                        // — one load instead of one per neighbour: 2.5-3.5 % slower, profiles/history/r02_watch_neighbour_ab.txt.)
                        // (Tried: also prefetching the records of every link that meets at the node — LinkStat, LinkHot,
                        // first queue sector — to break the chain action -> link records -> queue entries. Slower on B200:
                        // 2.81 vs 2.68 s at 1.02 M links, 221 vs 210 ms on config 2; profiles/history/r02_link_prefetch_ab.txt.)
                    }
                }
            }
            unsigned cm = __ballot_sync(0xffffffffu, cand);
#pragma unroll 1
            while (cm) {
                const int j = __ffs(cm) - 1;
                cm &= cm - 1;
                const int nn = __shfl_sync(0xffffffffu, my_node, j);
                const double my_t = __shfl_sync(0xffffffffu, kt, j);
                const double nb_bound = __shfl_sync(0xffffffffu, nbt, j);
                double nt;
#ifdef MZ_SIM_PROFILE
                const long long vc = clock64();
#endif
                const int st = async_visit<SIG>(d, nn, false, my_t, nb_bound, &nt, s_acc[warp]);
#ifdef MZ_SIM_PROFILE
                in_visit += clock64() - vc;
                ++n_calls;
#endif
                if ((int)lane == j) {
                    s_keyt[slot] = nt;
                    if (st == VISIT_DONE) s_node[slot] = -1;
                }
                if (st == VISIT_DONE) ++n_done;
                if (st != VISIT_WAIT) ran = true;
            }
            __syncwarp();
        }
        if (ran) {
            idle = 0;
        } else {
            if (++idle > max_idle_sweeps) {  // cannot happen: the node with the globally smallest key can always run
                if (lane == 0) d.counters[5] = 1;
                break;
            }
            __nanosleep(100);
        }
    }
    __syncwarp();
    if (lane < 5 && s_acc[warp][lane]) atomicAdd(d.counters + lane, s_acc[warp][lane]);
#ifdef MZ_SIM_PROFILE
    if (lane == 0 && n_own) {  // warp totals: cycles alive, cycles inside async_visit, sweeps, visit calls, warps
        atomicAdd(d.counters + 48, (unsigned long long)(clock64() - kc0));
        atomicAdd(d.counters + 49, (unsigned long long)in_visit);
        atomicAdd(d.counters + 50, (unsigned long long)n_sweeps);
        atomicAdd(d.counters + 51, (unsigned long long)n_calls);
        atomicAdd(d.counters + 52, 1ull);
    }
#endif
}

struct CudaBackend {
    // Device memory comes from a few large chunks (bump allocation): an engine has ~45 buffers and one cudaFree each
    // cost 0.2-0.7 s per mz_sim_destroy on B200 — more than the whole simulated hour.
    static constexpr size_t kChunkBytes = (size_t)256 << 20;
    struct Chunk {
        char *base;
        size_t size, used;
    };
    struct Ctx {
        cudaStream_t own = nullptr, stream = nullptr;
        cudaStream_t up = nullptr;  // uploads issued by the build's worker threads (h2d_worker)
        cudaEvent_t ev0 = nullptr, ev1 = nullptr;
        std::vector<Chunk> chunks;
        void *peer[MZ_MAXR] = {nullptr};
    };
    static int check_device(int device)
    {
        int ndev = 0;
        cudaError_t e = cudaGetDeviceCount(&ndev);
        if (e != cudaSuccess || ndev == 0) {
            cudaGetLastError();
            mz::set_error("no CUDA device available (%s); mezzo_b200 has no CPU fallback",
                          e == cudaSuccess ? "device count 0" : cudaGetErrorString(e));
            return MZ_ENODEV;
        }
        MZ_REQUIRE(device >= 0 && device < ndev, MZ_EINVAL, "device %d outside [0,%d)", device, ndev);
        MZ_CUDA(cudaSetDevice(device));
        return MZ_OK;
    }
    static int set_device(int device)
    {
        MZ_CUDA(cudaSetDevice(device));
        return MZ_OK;
    }
    static int ctx_create(Ctx &c)
    {
        MZ_CUDA(cudaStreamCreateWithFlags(&c.own, cudaStreamNonBlocking));
        MZ_CUDA(cudaStreamCreateWithFlags(&c.up, cudaStreamNonBlocking));
        MZ_CUDA(cudaEventCreate(&c.ev0));
        MZ_CUDA(cudaEventCreate(&c.ev1));
        c.stream = c.own;
        return MZ_OK;
    }
    static void ctx_destroy(Ctx &c)
    {
        if (c.ev0) cudaEventDestroy(c.ev0);
        if (c.ev1) cudaEventDestroy(c.ev1);
        if (c.own) cudaStreamDestroy(c.own);
        if (c.up) cudaStreamDestroy(c.up);
        for (Chunk &k : c.chunks) cudaFree(k.base);
        c.chunks.clear();
        cudaGetLastError();
    }
    static int alloc(Ctx &c, void **p, size_t bytes)
    {
        bytes = (bytes + 255) / 256 * 256;
        for (Chunk &k : c.chunks)
            if (k.size - k.used >= bytes) {
                *p = k.base + k.used;
                k.used += bytes;
                return MZ_OK;
            }
        Chunk k;
        k.size = std::max(bytes, kChunkBytes);
        k.used = bytes;
        void *q = nullptr;
        MZ_CUDA(cudaMalloc(&q, k.size));
        k.base = (char *)q;
        c.chunks.push_back(k);
        *p = q;
        return MZ_OK;
    }
    static void free(Ctx &, void *) {}  // released with the chunks in ctx_destroy
    // the arena is a plain cudaMalloc block: exportable with cudaIpcGetMemHandle, peer-mapped by the other ranks
    static int arena_alloc(Ctx &, void **p, size_t bytes, int)
    {
        MZ_CUDA(cudaMalloc(p, bytes));
        return MZ_OK;
    }
    static void arena_free(Ctx &, void *p, size_t) { cudaFree(p); }
    static void detach_peers(SimHost<CudaBackend> *s)
    {
        for (int r = 0; r < MZ_MAXR; ++r)
            if (s->ctx.peer[r]) {
                cudaIpcCloseMemHandle(s->ctx.peer[r]);
                s->ctx.peer[r] = nullptr;
            }
        cudaGetLastError();
    }
    static int h2d(Ctx &c, void *dst, const void *src, size_t bytes)
    {
        // pageable sources are staged before the call returns, so callers may pass temporaries
        MZ_CUDA(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyHostToDevice, c.stream));
        return MZ_OK;
    }
    // From a worker thread of the build, while the main thread is still assembling the other tables: a blocking copy on
    // the upload stream — the data is on the device when the call returns, so joining the thread orders it before
    // everything the main stream does afterwards.
    static int h2d_worker(Ctx &c, int device, void *dst, const void *src, size_t bytes)
    {
        if (!bytes) return MZ_OK;
        MZ_CUDA(cudaSetDevice(device));
        MZ_CUDA(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyHostToDevice, c.up));
        MZ_CUDA(cudaStreamSynchronize(c.up));
        return MZ_OK;
    }
    static int memset(Ctx &c, void *dst, int byte, size_t bytes)
    {
        MZ_CUDA(cudaMemsetAsync(dst, byte, bytes, c.stream));
        return MZ_OK;
    }
    static int d2d(Ctx &c, void *dst, const void *src, size_t bytes)
    {
        if (bytes) MZ_CUDA(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToDevice, c.stream));
        return MZ_OK;
    }
    static int fill64(Ctx &c, void *dst, unsigned long long pattern, size_t n, size_t stride)
    {
        if (n) k_fill64<<<(unsigned)std::min<size_t>((n + 255) / 256, 148 * 8), 256, 0, c.stream>>>((unsigned long long *)dst, pattern, n, stride);
        MZ_CUDA(cudaGetLastError());
        return MZ_OK;
    }
    static int sync(Ctx &c)
    {
        MZ_CUDA(cudaStreamSynchronize(c.stream));
        MZ_CUDA(cudaGetLastError());
        return MZ_OK;
    }
    static int d2h_sync(Ctx &c, void *dst, const void *src, size_t bytes)
    {
        MZ_CUDA(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToHost, c.stream));
        return sync(c);
    }
    static int launch_final(Ctx &c, const SimDev &d, int final_node)
    {
        k_sim_final<<<1, 32, 0, c.stream>>>(d, final_node);
        MZ_CUDA(cudaGetLastError());
        return MZ_OK;
    }
    static bool per_thread(const SimHost<CudaBackend> *s)
    {
        // Kernel choice: a warp per node (lane-parallel queue work, shortest latency per visit) while a block's slot table
        // holds its region, a thread per node beyond.
        return s->kernel_variant == 2 || (s->kernel_variant == 0 && (int)s->h_my_nodes.size() >= kThreadPerNodeFrom);
    }
    static bool notify(const SimHost<CudaBackend> *s) { return s->kernel_variant == 3; }
    // thread blocks (= regions) the node-visit kernel of the current variant runs with: every block resident, no more
    // blocks than there is work for
    static int default_regions(SimHost<CudaBackend> *s)
    {
        const int n_my = (int)s->h_my_nodes.size();
        int blocks = 0, per_block = 0;
        if (per_thread(s)) {
            if (mzgroup::max_blocks(s->device, &blocks) != MZ_OK) return 1;
            per_block = mzgroup::threads_per_block();
        } else {
            int per_sm = 0, sms = 0;
            // occupancy at the LARGEST slot table of either scheduler: a cooperative launch of that many blocks always fits
            const int smem_sweep = kMaxGroups * 32 * (kPThreads / 32) * kSweepSlotBytes, smem_notify = kMaxRegion * kSlotBytes + kMaxRegion / 32 * 8;
            const void *kerns[4] = {(const void *)k_sim_async<false>, (const void *)k_sim_async<true>, (const void *)k_sim_notify<false>,
                                    (const void *)k_sim_notify<true>};
            for (int k = 0; k < 4; ++k)
                if (cudaFuncSetAttribute(kerns[k], cudaFuncAttributeMaxDynamicSharedMemorySize, k < 2 ? smem_sweep : smem_notify) != cudaSuccess) {
                    cudaGetLastError();
                    return 1;
                }
            const bool nf = notify(s);
            if ((nf ? cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_sim_notify<true>, kPThreads, smem_notify)
                    : cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_sim_async<true>, kPThreads, smem_sweep)) != cudaSuccess ||
                cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, s->device) != cudaSuccess) {
                cudaGetLastError();
                return 1;
            }
            blocks = per_sm * sms;
            per_block = nf ? 4 * (kPThreads / 32) : kPThreads / 32;  // no more blocks than there is work for
        }
        blocks = std::max(1, std::min(blocks, (n_my + per_block - 1) / per_block));
        if (s->blocks_override > 0) blocks = std::min(blocks, s->blocks_override);
        return blocks;
    }
    // all node visits until no node of this rank has an event inside the horizon: one persistent kernel
    static int run_visits(SimHost<CudaBackend> *s, long long *steps, long long *launches)
    {
        Ctx &c = s->ctx;
        if (int rc = sim_set_regions(s, default_regions(s))) return rc;  // (no-op unless the variant / block count changed)
        SimDev d = s->dev;
        unsigned long long max_idle = 1ull << 28;  // ~30 s of fruitless polling: a bug, not a workload
        if (const char *mi = std::getenv("MZ_SIM_MAX_IDLE")) max_idle = std::strtoull(mi, nullptr, 10);  // debugging aid
        if (per_thread(s)) {
            if (int rc = mzgroup::launch(&d, sizeof(d), s->max_region, max_idle, c.stream)) return rc;
        } else if (notify(s)) {
            // one slot per node of the largest region (rounded up to whole mask words), then the READY and BUSY masks
            int cap = std::max(32, (s->max_region + 31) / 32 * 32);
            MZ_REQUIRE(cap <= kMaxRegion, MZ_ELIMIT,
                       "a region of %d nodes exceeds what one block tracks (%d); use more ranks or the thread-per-node kernel",
                       s->max_region, kMaxRegion);
            const size_t smem = (size_t)cap * kSlotBytes + (size_t)cap / 32 * 8;
            max_idle = 40ull * 1000 * 1000 * 1000;  // cycles without progress (~20 s): a bug, not a workload
            void *params[] = {&d, &max_idle, &cap};
            // cooperative launch only to guarantee that every block is resident (the nodes wait for each other)
            const void *kern = (d.sig_ev || d.trip_stop_off) ? (const void *)k_sim_notify<true> : (const void *)k_sim_notify<false>;
            MZ_CUDA(cudaLaunchCooperativeKernel(kern, dim3(s->n_regions), dim3(kPThreads), params, smem, c.stream));
        } else {
            // slots per warp: the region's nodes are dealt round-robin over the block's warps; an even number of 32-node groups
            const int warps = kPThreads / 32;
            int groups = (s->max_region + warps * 32 - 1) / (warps * 32);
            groups = std::max(2, (groups + 1) & ~1);
            MZ_REQUIRE(groups <= kMaxGroups, MZ_ELIMIT,
                       "a region of %d nodes exceeds what one block's warps track (%d); use more ranks or the thread-per-node kernel",
                       s->max_region, 32 * kMaxGroups * warps);
            const size_t smem = (size_t)groups * 32 * warps * kSweepSlotBytes;
            void *params[] = {&d, &max_idle, &groups};
            const void *kern = (d.sig_ev || d.trip_stop_off) ? (const void *)k_sim_async<true> : (const void *)k_sim_async<false>;
            MZ_CUDA(cudaLaunchCooperativeKernel(kern, dim3(s->n_regions), dim3(kPThreads), params, smem, c.stream));
        }
        MZ_CUDA(cudaStreamSynchronize(c.stream));
        MZ_CUDA(cudaGetLastError());
        ++*launches;
        *steps = 0;
        unsigned long long err = 0;
        MZ_CUDA(cudaMemcpy(&err, d.counters + 5, sizeof(err), cudaMemcpyDeviceToHost));
        MZ_REQUIRE(err != 3, MZ_ELIMIT, "a TurnAction of a signalised turning had more than %d live event chains", MZ_SIG_CHAINS);
        MZ_REQUIRE(err == 0, MZ_ECUDA, "link-update kernel gave up (code %llu): no node could run", err);
        return MZ_OK;
    }
    static int timer_start(Ctx &c)
    {
        MZ_CUDA(cudaEventRecord(c.ev0, c.stream));
        return MZ_OK;
    }
    static int timer_stop(Ctx &c, double *ms)
    {
        MZ_CUDA(cudaEventRecord(c.ev1, c.stream));
        MZ_CUDA(cudaStreamSynchronize(c.stream));
        MZ_CUDA(cudaGetLastError());
        float f = 0;
        MZ_CUDA(cudaEventElapsedTime(&f, c.ev0, c.ev1));
        *ms = f;
        return MZ_OK;
    }
};

}  // namespace

struct mz_sim : SimHost<CudaBackend> {};

extern "C" {

int mz_sim_create(int device, const MzNet *net, mz_sim **out)
{
    SimHost<CudaBackend> *s = nullptr;
    int rc = sim_create<CudaBackend>(device, net, SimPart{}, &s);
    if (out) *out = static_cast<mz_sim *>(s);
    return rc;
}

int mz_sim_create_part(int device, const MzNet *net, int n_ranks, int rank, const int32_t *node_rank, mz_sim **out)
{
    SimHost<CudaBackend> *s = nullptr;
    SimPart p;
    p.n_ranks = n_ranks;
    p.rank = rank;
    p.node_rank = node_rank;
    int rc = sim_create<CudaBackend>(device, net, p, &s);
    if (out) *out = static_cast<mz_sim *>(s);
    return rc;
}

int mz_sim_arena_handle(mz_sim *s, void *handle64, int64_t *arena_bytes)
{
    MZ_REQUIRE(s && handle64, MZ_EINVAL, "null argument");
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "IPC handle is 64 bytes");
    MZ_CUDA(cudaSetDevice(s->device));
    cudaIpcMemHandle_t h;
    MZ_CUDA(cudaIpcGetMemHandle(&h, s->arena));
    std::memcpy(handle64, &h, 64);
    if (arena_bytes) *arena_bytes = (int64_t)s->arena_bytes;
    return MZ_OK;
}

int mz_sim_attach_peer(mz_sim *s, int peer_rank, const void *handle64)
{
    MZ_REQUIRE(s && handle64, MZ_EINVAL, "null argument");
    MZ_REQUIRE(peer_rank >= 0 && peer_rank < s->n_ranks && peer_rank != s->rank, MZ_EINVAL, "bad peer rank %d", peer_rank);
    MZ_CUDA(cudaSetDevice(s->device));
    cudaIpcMemHandle_t h;
    std::memcpy(&h, handle64, 64);
    void *p = nullptr;
    MZ_CUDA(cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess));
    s->ctx.peer[peer_rank] = p;
    s->dev.arena[peer_rank] = (char *)p;
    bool all = true;
    for (int r = 0; r < s->n_ranks; ++r) all = all && s->dev.arena[r] != nullptr;
    s->peers_attached = all;
    return MZ_OK;
}

// Peers inside ONE process (a single mezzo_lib process driving several GPUs, host/NetworkStep.h): no IPC handle — CUDA cannot
// open its own process's handles — but direct peer access between the two devices and the peer's arena pointer as it is.
int mz_sim_attach_peer_local(mz_sim *s, int peer_rank, mz_sim *peer)
{
    MZ_REQUIRE(s && peer && s != peer, MZ_EINVAL, "null or identical engines");
    MZ_REQUIRE(peer_rank >= 0 && peer_rank < s->n_ranks && peer_rank != s->rank && peer->rank == peer_rank && peer->n_ranks == s->n_ranks,
               MZ_EINVAL, "bad peer rank %d", peer_rank);
    MZ_REQUIRE(peer->arena_bytes == s->arena_bytes, MZ_EINVAL, "the two engines were not created from the same network and partition");
    MZ_CUDA(cudaSetDevice(s->device));
    if (peer->device != s->device) {
        int can = 0;
        MZ_CUDA(cudaDeviceCanAccessPeer(&can, s->device, peer->device));
        MZ_REQUIRE(can, MZ_ECUDA, "device %d cannot access device %d (no NVLink / P2P path)", s->device, peer->device);
        const cudaError_t e = cudaDeviceEnablePeerAccess(peer->device, 0);
        if (e == cudaErrorPeerAccessAlreadyEnabled) cudaGetLastError();
        else MZ_CUDA(e);
    }
    s->dev.arena[peer_rank] = (char *)peer->arena;  // (not in ctx.peer: nothing to close on destroy)
    bool all = true;
    for (int r = 0; r < s->n_ranks; ++r) all = all && s->dev.arena[r] != nullptr;
    s->peers_attached = all;
    return MZ_OK;
}

int mz_partition_nodes(const MzNet *net, int n_ranks, int32_t *node_rank) { return partition_nodes(net, n_ranks, node_rank); }

int mz_sim_reset(mz_sim *s)
{
    MZ_REQUIRE(s, MZ_EINVAL, "null handle");
    return sim_reset_state<CudaBackend>(s);
}

int mz_sim_run(mz_sim *s) { return sim_run<CudaBackend>(s); }

int mz_sim_min_key(mz_sim *s, double *t, double *tp, int32_t *sub, int32_t *node)
{
    MZ_REQUIRE(s && t && tp && sub && node, MZ_EINVAL, "null argument");
    Key k;
    if (int rc = sim_min_key<CudaBackend>(s, &k)) return rc;
    *t = k.t;
    *tp = k.tp;
    *sub = k.sub;
    *node = k.id == 0x7fffffff ? -1 : k.id;
    return MZ_OK;
}

int mz_sim_final_event(mz_sim *s, int32_t final_node, double end_time)
{
    return sim_final_event<CudaBackend>(s, final_node, end_time);
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
    return sim_get_exit_log<CudaBackend>(s, out, cap, n_out);
}

int mz_sim_get_linktimes(mz_sim *s, double *avg)
{
    MZ_REQUIRE(s && avg, MZ_EINVAL, "null argument");
    MZ_CUDA(cudaSetDevice(s->device));
    MZ_CUDA(cudaMemcpy(avg, s->dev.avgtimes, sizeof(double) * (size_t)s->n_links * s->n_periods, cudaMemcpyDefault));
    return MZ_OK;
}

int mz_sim_get_linktimes_final(mz_sim *s, double *avg) { return sim_get_linktimes_final<CudaBackend>(s, avg); }

int mz_sim_last_stats(const mz_sim *s, mz_sim_stats *out)
{
    MZ_REQUIRE(s && out, MZ_EINVAL, "null argument");
    *out = s->stats;
    return MZ_OK;
}

int mz_sim_set_stream(mz_sim *s, void *cuda_stream)
{
    MZ_REQUIRE(s, MZ_EINVAL, "null handle");
    s->ctx.stream = cuda_stream ? (cudaStream_t)cuda_stream : s->ctx.own;
    return MZ_OK;
}

int mz_sim_set_option(mz_sim *s, int option, int64_t value)
{
    MZ_REQUIRE(s, MZ_EINVAL, "null handle");
    switch (option) {
        case MZ_SIM_OPT_MAX_EVENTS:
            MZ_REQUIRE(value >= 1, MZ_EINVAL, "max events per window must be >= 1");
            s->dev.max_events = (int)std::min<int64_t>(value, 1 << 30);
            return MZ_OK;
        case MZ_SIM_OPT_BLOCKS:
            MZ_REQUIRE(value >= 0, MZ_EINVAL, "block count must be >= 0 (0 = automatic)");
            s->blocks_override = (int)value;
            return MZ_OK;
        case MZ_SIM_OPT_KERNEL:
            MZ_REQUIRE(value >= 0 && value <= 3, MZ_EINVAL,
                       "kernel variant must be 0 (automatic), 1 (warp per node, sweeps), 2 (thread per node) or 3 (warp per node, wake-ups)");
            s->kernel_variant = (int)value;
            return MZ_OK;
        case MZ_SIM_OPT_CHECK_EVERY:
            return MZ_OK;  // kept for ABI compatibility: the run is one kernel, the host never polls
        default:
            mz::set_error("unknown option %d", option);
            return MZ_EINVAL;
    }
}

/* diagnostics (not part of the drop-in surface): per-event-kind cycle counters of the last run; all zero unless the
 * library was built with -DMZ_SIM_PROFILE (tools/sim_profile.py) */
int mz_sim_debug_prof_events(const mz_sim *s, uint64_t out[56])
{
    MZ_REQUIRE(s && out, MZ_EINVAL, "null argument");
    for (int i = 0; i < 56; ++i) out[i] = s->prof_ev[i];
    return MZ_OK;
}

int mz_sim_run_until(mz_sim *s, double t) { return sim_run_until<CudaBackend>(s, t); }
int mz_sim_stop_arrivals(mz_sim *s, mz_stop_visit *out, int64_t cap, int64_t *n_out) { return sim_stop_arrivals<CudaBackend>(s, out, cap, n_out); }
int mz_sim_stop_departures(mz_sim *s, int64_t n, const int32_t *veh, const double *t_depart)
{
    return sim_stop_departures<CudaBackend>(s, n, veh, t_depart);
}
int mz_sim_destroy(mz_sim *s) { return sim_destroy<CudaBackend>(s); }
}
