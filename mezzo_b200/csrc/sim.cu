// Hot path 1: the link update on sm_100a — CUDA backend + C ABI (mz_sim_*).
// The data layout and per-event functions live in sim_core.h, the backend-independent host logic in sim_host.h.
#include "common.h"
#include "sim_host.h"

using namespace mzsim;

namespace {

// One conservative window per launch: one WARP per node. Used for the clean-up pass that executes exactly one event
// (`final_node` >= 0: the first event at/after the horizon) and as a fallback driver.
constexpr int kSimThreads = 128;
__global__ void __launch_bounds__(kSimThreads) k_sim_superstep(SimDev d, int parity, int final_node)
{
    const int i = (int)((blockIdx.x * (unsigned)blockDim.x + threadIdx.x) >> 5);
    if (i >= d.n_my) return;  // warp-uniform
    const bool alive = superstep_node(d, lds(d.my_nodes + i), parity, final_node);
    if (alive && (threadIdx.x & 31) == 0) d.counters[4] = 1;  // somebody still has work inside the horizon
}

// ---- the persistent window loop -----------------------------------------------------------------------------------
// All windows of a run execute inside ONE cooperative kernel: a warp owns a fixed, strided set of this rank's nodes
// (neighbouring nodes sit on different warps), static tables stay L1-resident, and a window ends with a grid barrier
// instead of a kernel boundary. With several ranks the last block to arrive also exchanges (window, alive) words with
// the peer GPUs through their arenas (NVLink stores + local polling) before it releases the local blocks.
struct PCtl {
    unsigned *bar;                  // [0] arrival count  [1] generation  [2..4] alive flags (by window % 3)  [5] windows done  [6] finished
    unsigned long long *sync_base;  // this rank's arena sync words: [r] = last window rank r completed (+ alive bit)
    size_t off_sync;                // offset of the sync words inside every arena
    int parity0;
    unsigned max_windows;
    unsigned long long win0;        // global index of the first window of this launch
};
#ifndef MZ_PTHREADS
#define MZ_PTHREADS 512
#endif
constexpr int kPThreads = MZ_PTHREADS;

__device__ __forceinline__ bool window_barrier(const SimDev &d, const PCtl &c, unsigned long long win, bool block_alive)
{
    const unsigned slot = (unsigned)(win % 3);
    __syncthreads();
    if (threadIdx.x == 0) {
        volatile unsigned *bar = c.bar;
        if (block_alive) bar[2 + slot] = 1;
        const unsigned g = bar[1];
        __threadfence();
        if (atomicAdd(c.bar, 1u) == gridDim.x - 1) {
            // last block of this rank: everything this rank did in the window is visible (each block fenced before arriving)
            if (d.n_ranks > 1) {
                __threadfence_system();
                const unsigned long long word = ((win + 1) << 1) | (bar[2 + slot] ? 1ull : 0ull);
                for (int r = 0; r < d.n_ranks; ++r) {
                    volatile unsigned long long *w =
                        reinterpret_cast<volatile unsigned long long *>(d.arena[r] + c.off_sync) + d.rank;
                    *w = word;
                }
                unsigned any = 0;
                for (int r = 0; r < d.n_ranks; ++r) {
                    volatile unsigned long long *w = c.sync_base + r;
                    unsigned long long v;
                    while (((v = *w) >> 1) < win + 1) __nanosleep(40);
                    any |= (unsigned)(v & 1ull);
                }
                __threadfence_system();
                if (any) bar[2 + slot] = 1;
            }
            bar[2 + (slot + 1) % 3] = 0;  // next window's flag: last read two barriers ago
            bar[0] = 0;
            __threadfence();
            atomicAdd(c.bar + 1, 1u);
        } else {
            while (bar[1] == g) __nanosleep(20);
        }
        __threadfence();
    }
    __syncthreads();
    return reinterpret_cast<volatile unsigned *>(c.bar)[2 + slot] != 0;
}

constexpr int kMaxCand = 1024;
__global__ void __launch_bounds__(kPThreads, 1) k_sim_persistent(SimDev d, PCtl c)
{
    __shared__ int s_cand[kMaxCand];
    __shared__ unsigned s_ncand, s_next;
    const unsigned lane = threadIdx.x & 31;
    int parity = c.parity0;
    unsigned w = 0;
    bool go = true;
    if (threadIdx.x == 0) {
        s_ncand = 0;
        s_next = 0;
    }
    __syncthreads();
    long long prof[3] = {0, 0, 0};  // block 0's view: cycles in phase 1 / phase 2 / barrier
    for (; w < c.max_windows && go; ++w) {
        bool alive = false;
        const long long c0 = clock64();
        // Phase 1 — one THREAD per node (block b owns nodes b, b+grid, ...: neighbouring nodes sit on different SMs):
        // own key against the neighbours' times. Nodes that cannot be a local minimum just re-publish their key; the
        // candidates go on the block's list.
        for (unsigned base = blockIdx.x; base < (unsigned)d.n_my; base += gridDim.x * kPThreads) {
            const unsigned idx = base + threadIdx.x * gridDim.x;
            if (idx < (unsigned)d.n_my) {
                const int n = lds(d.my_nodes + idx);
                const Key mine = node_key(d, n, parity);
                bool cand = false;
                if (mine.t < d.runtime) {
                    alive = true;
                    cand = true;
                    // eight neighbour keys in flight at a time (a junction has 4-8 neighbours): one load round, not one per neighbour
                    const int nb0 = lds(d.node_nb_off + n), nb1 = lds(d.node_nb_off + n + 1);
                    for (int i0 = nb0; i0 < nb1; i0 += 8) {
                        double mt[8];
#pragma unroll
                        for (int k = 0; k < 8; ++k) {
                            mt[k] = INFINITY;
                            if (i0 + k < nb1) {
                                const int m = lds(d.node_nb + i0 + k);
                                mt[k] = ldm(reinterpret_cast<const double *>(d.arena[node_rank(d, m)] + d.off_kt[parity]) + m);
                            }
                        }
#pragma unroll
                        for (int k = 0; k < 8; ++k)
                            if (mt[k] < mine.t) cand = false;
                    }
                }
                unsigned slot = kMaxCand;
                if (cand) slot = atomicAdd(&s_ncand, 1u);
                if (slot < kMaxCand) {
                    s_cand[slot] = n;
                } else if (cand) {
                    // list full (never on road networks): the node simply waits for the next window
                    cand = false;
                }
                if (!cand) {
                    char *a = d.arena[d.rank];
                    reinterpret_cast<double *>(a + d.off_kt[parity ^ 1])[n] = mine.t;
                    reinterpret_cast<double *>(a + d.off_ktp[parity ^ 1])[n] = mine.tp;
                    reinterpret_cast<int *>(a + d.off_ksub[parity ^ 1])[n] = mine.sub;
                }
            }
        }
        __syncthreads();
        const long long c1 = clock64();
        // Phase 2 — one WARP per candidate, taken dynamically so no warp runs two windows while others idle
        const unsigned nc = s_ncand < kMaxCand ? s_ncand : kMaxCand;
        for (;;) {
            unsigned i = 0;
            if (lane == 0) i = atomicAdd(&s_next, 1u);
            i = __shfl_sync(0xffffffffu, i, 0);
            if (i >= nc) break;
            superstep_node(d, s_cand[i], parity, -1);  // re-publishes the node's key itself
        }
        const bool any = __syncthreads_or(alive);
        const long long c2 = clock64();
        if (threadIdx.x == 0) {
            s_ncand = 0;
            s_next = 0;
        }
        go = window_barrier(d, c, c.win0 + w, any);
        parity ^= 1;
        const long long c3 = clock64();
        prof[0] += c1 - c0;
        prof[1] += c2 - c1;
        prof[2] += c3 - c2;
    }
    if (blockIdx.x == 0 && threadIdx.x == 0) {
        c.bar[5] = w;
        c.bar[6] = go ? 0u : 1u;
        d.counters[5] += (unsigned long long)prof[0];
        d.counters[6] += (unsigned long long)prof[1];
        d.counters[7] += (unsigned long long)prof[2];
    }
}

struct CudaBackend {
    struct Ctx {
        cudaStream_t own = nullptr, stream = nullptr;
        cudaEvent_t ev0 = nullptr, ev1 = nullptr;
        unsigned *bar = nullptr;
        int blocks = 0;
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
        MZ_CUDA(cudaEventCreate(&c.ev0));
        MZ_CUDA(cudaEventCreate(&c.ev1));
        MZ_CUDA(cudaMalloc((void **)&c.bar, 16 * sizeof(unsigned)));
        MZ_CUDA(cudaMemset(c.bar, 0, 16 * sizeof(unsigned)));
        c.stream = c.own;
        return MZ_OK;
    }
    static void ctx_destroy(Ctx &c)
    {
        if (c.ev0) cudaEventDestroy(c.ev0);
        if (c.ev1) cudaEventDestroy(c.ev1);
        if (c.own) cudaStreamDestroy(c.own);
        if (c.bar) cudaFree(c.bar);
        cudaGetLastError();
    }
    static int alloc(void **p, size_t bytes)
    {
        MZ_CUDA(cudaMalloc(p, bytes));
        return MZ_OK;
    }
    static void free(void *p) { cudaFree(p); }
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
    static int memset(Ctx &c, void *dst, int byte, size_t bytes)
    {
        MZ_CUDA(cudaMemsetAsync(dst, byte, bytes, c.stream));
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
    static int launch(Ctx &c, const SimDev &d, int parity, int final_node)
    {
        const int nodes_per_block = kSimThreads / 32;
        k_sim_superstep<<<(d.n_my + nodes_per_block - 1) / nodes_per_block, kSimThreads, 0, c.stream>>>(d, parity,
                                                                                                      final_node);
        MZ_CUDA(cudaGetLastError());
        return MZ_OK;
    }
    // all windows until no rank has an event inside the horizon: chunks of the persistent kernel
    static int run_windows(SimHost<CudaBackend> *s, long long *steps, long long *launches)
    {
        Ctx &c = s->ctx;
        if (c.blocks == 0) {
            int per_sm = 0, sms = 0;
            MZ_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_sim_persistent, kPThreads, 0));
            MZ_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, s->device));
            MZ_REQUIRE(per_sm > 0, MZ_ECUDA, "k_sim_persistent does not fit on an SM");
            // no more warps than nodes: idle blocks only lengthen the barrier
            const int want = (s->dev.n_my + (kPThreads / 32) - 1) / (kPThreads / 32);
            c.blocks = std::max(1, std::min(per_sm * sms, want));
            if (s->blocks_override > 0) c.blocks = std::min(per_sm * sms, s->blocks_override);
        }
        MZ_CUDA(cudaMemsetAsync(c.bar, 0, 16 * sizeof(unsigned), c.stream));
        unsigned long long win = 0;
        for (;;) {
            PCtl pc;
            pc.bar = c.bar;
            pc.sync_base = reinterpret_cast<unsigned long long *>((char *)s->arena + s->off_sync);
            pc.off_sync = s->off_sync;
            pc.parity0 = s->parity;
            pc.max_windows = (unsigned)s->windows_per_launch;
            pc.win0 = win;
            SimDev d = s->dev;
            void *params[] = {&d, &pc};
            MZ_CUDA(cudaLaunchCooperativeKernel((const void *)k_sim_persistent, dim3(c.blocks), dim3(kPThreads), params, 0,
                                                c.stream));
            unsigned h[8];
            MZ_CUDA(cudaMemcpyAsync(h, c.bar, sizeof(h), cudaMemcpyDeviceToHost, c.stream));
            MZ_CUDA(cudaStreamSynchronize(c.stream));
            MZ_CUDA(cudaGetLastError());
            ++*launches;
            *steps += h[5];
            win += h[5];
            if (h[5] & 1u) s->parity ^= 1;
            if (h[6]) break;
            MZ_REQUIRE(win < (1ull << 40), MZ_ECUDA, "window loop does not terminate");
        }
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
            s->ctx.blocks = 0;
            return MZ_OK;
        case MZ_SIM_OPT_CHECK_EVERY:
            MZ_REQUIRE(value >= 1 && value <= (1 << 24), MZ_EINVAL, "windows per launch must be in [1,2^24]");
            s->windows_per_launch = (int)value;
            return MZ_OK;
        default:
            mz::set_error("unknown option %d", option);
            return MZ_EINVAL;
    }
}

/* diagnostics (not part of the drop-in surface): block 0's cycles in phase 1 / phase 2 / barrier of the last run */
int mz_sim_debug_prof(const mz_sim *s, uint64_t out[3])
{
    MZ_REQUIRE(s && out, MZ_EINVAL, "null argument");
    for (int i = 0; i < 3; ++i) out[i] = s->prof[i];
    return MZ_OK;
}
int mz_sim_debug_prof_events(const mz_sim *s, uint64_t out[24])
{
    MZ_REQUIRE(s && out, MZ_EINVAL, "null argument");
    for (int i = 0; i < 24; ++i) out[i] = s->prof_ev[i];
    return MZ_OK;
}

int mz_sim_destroy(mz_sim *s) { return sim_destroy<CudaBackend>(s); }
}
