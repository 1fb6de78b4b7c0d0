// Hot path 2, throughput kernels: frontier-based near-far (delta-stepping style) label-correcting over a whole chunk
// of trees at once (MZ_SSSP_AUTO).
//
// Reference: Graph::labelCorrecting B/Graph.cpp:313-454, LinkTimeInfo::cost B/linktimes.cpp:108-121
// (B/ = /root/reference/branches/busmezzo/mezzo_lib/src/).
//
// Why this may replace the sequential deque and still be bit-exact: with a link-time table whose rows are
// non-negative and non-decreasing in the period index ("FIFO table"), f_v(x) = x (+) T[v][int((entry+x)/periodlength)] is
// monotone in x under IEEE rounding, so EVERY label-correcting order ends at the same least fix-point of labels; the
// predecessor the reference keeps is the in-neighbour whose relaxation produced the final label, which is unique whenever
// exactly one in-neighbour u satisfies label[u] (+) cost == label[v].  k_fast_finalize checks this per tree on the device
// (fix-point, unique argmin, every cost lookup inside the table, FIFO rows); a tree that fails any check — ties, non-FIFO
// rows, labels running past the end of the table — is recomputed by the exact deque kernel (k_sssp_exact).  The graph
// must have identical successor sets for all in-links of a node (no turning prohibitors), otherwise the reference's
// literal "queue[rear] = v" overwrite can lose list entries and only the exact kernel reproduces it.
//
// Formulation: the chunk's n_trees x L (tree, link) states form ONE label-correcting problem. Work item = one state,
// one THREAD per item (32 independent dependent-load chains per warp); a single threshold and one near / one far list
// serve all trees, since every tree's labels start at ~0 and grow on the same scale. One persistent cooperative kernel
// runs all steps with a hand-rolled grid barrier (one barrier per step):
//   step s: near list [s%3] is consumed, improved states with label < thr are pushed to [(s+1)%3], the others to the far
//           list. Pushes are NOT de-duplicated (that cost one more dependent atomic after the atomicMin, 27 % of the
//           stall samples in ncu): a state improved twice in a step is listed twice and the second copy is dropped when
//           it is popped (atomicExch on its stamp, issued together with the label load);
//   when the near list is empty the threshold advances (thr = max(thr, min far label) + delta) and the far list is split:
//           label < old thr: already scanned via the near path, dropped; < new thr: to the near list; else kept.
// (A first version used lanes = trees with label rows [link][32]; with the 2 entry times per root of config 2 only 2 of
// 32 lanes shared a frontier link and the kernel ran at 1 ms per step — item-rate bound, not bandwidth bound.)
#pragma once
#include <cuda_runtime.h>

namespace mzfast {

constexpr int RUN_BLOCK = 256;
constexpr int RUN_BLOCKS_PER_SM = 4;
constexpr double INF_LABEL = MZ_SP_INFINITY;

struct Ctl {  // triple-buffered by step (near) / by advance count (far)
    unsigned ncnt[3];
    unsigned fcnt[3];
    unsigned long long farmin[3];  // bit pattern of the smallest label in the far list (lower bound)
    unsigned pad[2];
};

// per (tree, link) state: label, near-list stamp and far-list flag share one 16-byte record, so the relaxation's read,
// its atomicMin and the list bookkeeping touch ONE 32-byte sector (three separate arrays cost three DRAM misses)
struct alignas(16) State {
    unsigned long long lab;  // bit pattern of the f64 label (non-negative doubles order like their bits)
    unsigned stamp;          // step (+1) in which the state was last scanned: drops duplicate list entries at POP time
    unsigned pad;
};
static_assert(sizeof(State) == 16, "State must be 16 bytes");

struct RunArgs {
    const int *succ_off;
    const int2 *succ;  // {v, dnNode(v)}
    const double *T;  // period-major copy of the table (Tt[period][link])
    int P;
    double periodlength;
    unsigned L, nt;
    const int *rootnode;  // [nt] up-node of each tree's root link
    const double *entry;  // [nt]
    State *st;            // [nt][L]
    unsigned *nearl;      // [3][cap] state indices t*L+v
    unsigned *farl;       // [3][cap]
    unsigned cap;
    Ctl *ctl;
    unsigned *bar;  // [0] arrival count, [1] generation
    unsigned long long *evaluated;
    unsigned *status;  // [0] aborted (step cap or list overflow), [1] steps executed
    double delta;
    unsigned max_steps;
};

// The frontier kernels read the link-time table PERIOD-MAJOR (Tt[period][link]): all successors of a scanned link are looked
// up in the same period (it follows from the scanned link's label) and have consecutive ids, so their costs share one
// 64-byte burst — with the reference's link-major rows (128 B per link at P = 16, a 320 MB table on config 4) every lookup
// was a burst of its own.
__device__ __forceinline__ double fast_cost(const double *__restrict__ Tt, int P, unsigned L, int idx, unsigned l)
{
    if (idx < 0 || idx >= P) return MZ_MISSING_COST;
    return __ldg(Tt + (size_t)idx * L + l);
}
__global__ void k_transpose_table(const double *__restrict__ T, double *__restrict__ Tt, int P, unsigned L)
{
    const size_t n = (size_t)P * L, stride = (size_t)gridDim.x * blockDim.x;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        const size_t l = i / (size_t)P, p = i - l * (size_t)P;
        Tt[p * L + l] = T[i];
    }
}

__device__ __forceinline__ void grid_barrier(unsigned *bar, unsigned nblocks)
{
    __syncthreads();
    if (threadIdx.x == 0) {
        volatile unsigned *gen = bar + 1;
        const unsigned g = *gen;
        __threadfence();
        if (atomicAdd(bar, 1u) == nblocks - 1) {
            bar[0] = 0;
            __threadfence();
            atomicAdd(bar + 1, 1u);
        } else {
            while (*gen == g) __nanosleep(20);
        }
        __threadfence();
    }
    __syncthreads();
}

// minimum over the lanes in `mask` of a non-negative double (bit patterns of non-negative doubles order like integers)
__device__ __forceinline__ unsigned long long warp_min_bits(unsigned mask, unsigned long long bits)
{
    const unsigned hi = (unsigned)(bits >> 32), lo = (unsigned)bits;
    const unsigned mh = __reduce_min_sync(mask, hi);
    const unsigned ml = __reduce_min_sync(mask, hi == mh ? lo : 0xffffffffu);
    return ((unsigned long long)mh << 32) | ml;
}

// Warp-level staging of list pushes: each warp appends to its own shared-memory buffer (positions from a ballot, no
// atomics) and, when the buffer runs full or the step ends, reserves a range of the global list with ONE global atomicAdd
// and copies the staged items out coalesced.  (Pushing straight to the global counters — three same-address atomics per
// warp iteration — serialised in one L2 slice and held the kernel at ~4 G items/s; block-level staging with
// __syncthreads per round left 28 % of the issue slots stalled on the barrier.)
constexpr int WSTAGE = 512;        // entries per warp and list
constexpr int UNROLL = 4;          // successors relaxed together (independent loads/atomics in flight per thread)
struct WStage {
    unsigned *buf;  // this warp's shared-memory buffer
    unsigned n;     // warp-uniform fill
};
__device__ __forceinline__ void wstage_flush(WStage &w, unsigned *list, unsigned *cnt, unsigned cap)
{
    if (w.n == 0) return;
    const int lane = threadIdx.x & 31;
    unsigned base = 0;
    if (lane == 0) base = atomicAdd(cnt, w.n);
    base = __shfl_sync(0xffffffffu, base, 0);
    __syncwarp();
    // beyond cap: dropped; the global counter still counts them and the next step aborts
    for (unsigned i = lane; i < w.n; i += 32)
        if (base + i < cap) list[base + i] = w.buf[i];
    __syncwarp();
    w.n = 0;
}
// all 32 lanes converged; `want` selects the pushers
__device__ __forceinline__ void wstage_push(bool want, unsigned item, WStage &w)
{
    const unsigned m = __ballot_sync(0xffffffffu, want);
    if (want) w.buf[w.n + __popc(m & ((1u << (threadIdx.x & 31)) - 1u))] = item;
    w.n += __popc(m);
}

__global__ void k_fast_init(State *__restrict__ st, size_t n)
{
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    const unsigned long long inf = (unsigned long long)__double_as_longlong(INF_LABEL);
    uint4 v;
    v.x = (unsigned)inf;
    v.y = (unsigned)(inf >> 32);
    v.z = 0;
    v.w = 0;
    uint4 *p = reinterpret_cast<uint4 *>(st);
    for (size_t j = (size_t)blockIdx.x * blockDim.x + threadIdx.x; j < n; j += stride) p[j] = v;
}

// label[root] = cost(root, entry); the root state goes on the first near list (B/Graph.cpp:356-363)
__global__ void k_fast_seed(RunArgs a, const int *__restrict__ rootlink)
{
    const unsigned t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= a.nt) return;
    const unsigned s = (unsigned)rootlink[t];
    const int idx = (int)(a.entry[t] / a.periodlength);
    const unsigned st = t * a.L + s;
    a.st[st].lab = (unsigned long long)__double_as_longlong(fast_cost(a.T, a.P, a.L, idx, s));
    a.nearl[t] = st;  // near list 0, one entry per tree (cap >= nt)
}
__global__ void k_fast_ctl_init(RunArgs a)
{
    Ctl c;
    for (int k = 0; k < 3; ++k) {
        c.ncnt[k] = 0;
        c.fcnt[k] = 0;
        c.farmin[k] = (unsigned long long)__double_as_longlong(INFINITY);
    }
    c.ncnt[0] = a.nt;
    c.pad[0] = c.pad[1] = 0;
    *a.ctl = c;
    a.bar[0] = 0;
    a.bar[1] = 0;
    a.status[0] = 0;
    a.status[1] = 0;
}

__global__ void __launch_bounds__(RUN_BLOCK, RUN_BLOCKS_PER_SM) k_fast_run(RunArgs a)
{
    __shared__ unsigned s_stage[2][RUN_BLOCK / 32][WSTAGE];
    __shared__ unsigned long long s_fmin;
    const unsigned tid = blockIdx.x * RUN_BLOCK + threadIdx.x, nthreads = gridDim.x * RUN_BLOCK;
    const unsigned L = a.L;
    Ctl *c = a.ctl;
    double thr = a.delta;
    int fp = 0;
    bool advprev = false;
    unsigned long long n_eval = 0;
    unsigned s = 0;
    WStage w_near{s_stage[0][threadIdx.x >> 5], 0}, w_far{s_stage[1][threadIdx.x >> 5], 0};
    if (threadIdx.x == 0) s_fmin = ~0ull;
    __syncthreads();
    for (;; ++s) {
        const int cur = s % 3, nxt = (s + 1) % 3, old = (s + 2) % 3;
        // ---- decision: identical in every thread (reads only words nobody writes during this step)
        if (tid == 0) {  // housekeeping of slots nobody touches in this step
            if (advprev) {
                const int dead = (fp + 2) % 3;  // far slot consumed by the previous step's split
                __stcg(&c->fcnt[dead], 0u);
                __stcg(&c->farmin[dead], (unsigned long long)__double_as_longlong(INFINITY));
            }
            __stcg(&c->ncnt[old], 0u);
        }
        advprev = false;
        unsigned items = __ldcg(&c->ncnt[cur]);
        int mode = 1;
        const double thr_old = thr;
        if (items == 0) {
            items = __ldcg(&c->fcnt[fp]);
            mode = 2;
            if (items) {
                const double fm = __longlong_as_double((long long)__ldcg(&c->farmin[fp]));
                thr = (fm > thr ? fm : thr) + a.delta;
            }
        }
        if (items == 0) break;
        // a list that outgrew its capacity lost entries (the counters still count them): give up, the host redoes the
        // chunk with the exact kernel. Both tests read words that are stable during this step, so every thread agrees.
        if (items > a.cap || s >= a.max_steps) {
            if (tid == 0) a.status[0] = 1;
            break;
        }
        unsigned *near_nxt = a.nearl + (size_t)nxt * a.cap;
        unsigned long long fmin = ~0ull;  // smallest label this thread sent to the far list
        const unsigned rounds = (items + nthreads - 1) / nthreads;
        int far_slot = fp;
        if (mode == 1) {
            const unsigned *near_cur = a.nearl + (size_t)cur * a.cap;
            unsigned *far_cur = a.farl + (size_t)fp * a.cap;
            for (unsigned r = 0; r < rounds; ++r) {
                // consecutive items to consecutive blocks' threads: a block's items are contiguous in the list
                const unsigned i = (r * gridDim.x + blockIdx.x) * RUN_BLOCK + threadIdx.x;
                unsigned t = 0, sb = 0, se = 0;
                double lu = 0.0;
                int rootnode = -1, pidx = 0;
                if (i < items) {
                    const unsigned st = __ldcg(near_cur + i);
                    t = st / L;
                    const unsigned u = st - t * L;
                    const bool dup = atomicExch(&a.st[st].stamp, s + 1u) == s + 1u;  // already scanned in this step
                    lu = __longlong_as_double((long long)__ldcg(&a.st[st].lab));
                    if (!dup) {
                        sb = (unsigned)__ldg(a.succ_off + u);
                        se = (unsigned)__ldg(a.succ_off + u + 1);
                    }
                    rootnode = __ldg(a.rootnode + t);
                    pidx = (int)((__ldg(a.entry + t) + lu) / a.periodlength);  // B/linktimes.cpp:63, same for all successors
                }
                for (unsigned jb = sb; __any_sync(0xffffffffu, jb < se); jb += UNROLL) {
                    // four successors at a time: their loads and atomics are independent, so they overlap
                    unsigned vs[UNROLL];
                    unsigned long long nlb[UNROLL], prev[UNROLL];
                    bool cand[UNROLL];
#pragma unroll
                    for (int k = 0; k < UNROLL; ++k) {
                        cand[k] = false;
                        vs[k] = 0;
                        nlb[k] = 0;
                        if (jb + k < se) {
                            const int2 sv = __ldg(a.succ + jb + k);
                            if (sv.y != rootnode) {  // B/Graph.cpp:385
                                const double nl = lu + fast_cost(a.T, a.P, L, pidx, (unsigned)sv.x);  // B/Graph.cpp:399
                                ++n_eval;
                                vs[k] = t * L + (unsigned)sv.x;
                                nlb[k] = (unsigned long long)__double_as_longlong(nl);
                                cand[k] = true;
                            }
                        }
                    }
#pragma unroll
                    for (int k = 0; k < UNROLL; ++k)  // strict '<', B/Graph.cpp:412 (non-negative doubles order like their bits)
                        if (cand[k]) cand[k] = nlb[k] < __ldcg(&a.st[vs[k]].lab);
#pragma unroll
                    for (int k = 0; k < UNROLL; ++k) {
                        prev[k] = 0;
                        if (cand[k]) prev[k] = atomicMin(&a.st[vs[k]].lab, nlb[k]);
                    }
                    bool to_near[UNROLL], to_far[UNROLL];
#pragma unroll
                    for (int k = 0; k < UNROLL; ++k) {
                        to_near[k] = to_far[k] = false;
                        if (cand[k] && nlb[k] < prev[k]) {
                            if (__longlong_as_double((long long)nlb[k]) < thr) {
                                to_near[k] = true;
                            } else {
                                to_far[k] = true;
                                fmin = nlb[k] < fmin ? nlb[k] : fmin;
                            }
                        }
                    }
#pragma unroll
                    for (int k = 0; k < UNROLL; ++k) {
                        wstage_push(to_near[k], vs[k], w_near);
                        wstage_push(to_far[k], vs[k], w_far);
                    }
                    if (w_near.n > WSTAGE - 32 * UNROLL) wstage_flush(w_near, near_nxt, &c->ncnt[nxt], a.cap);
                    if (w_far.n > WSTAGE - 32 * UNROLL) wstage_flush(w_far, far_cur, &c->fcnt[fp], a.cap);
                }
            }
            wstage_flush(w_near, near_nxt, &c->ncnt[nxt], a.cap);
            wstage_flush(w_far, far_cur, &c->fcnt[fp], a.cap);
        } else {  // split the far list against the advanced threshold
            const int fn = (fp + 1) % 3;
            const unsigned *far_cur = a.farl + (size_t)fp * a.cap;
            unsigned *far_nxt = a.farl + (size_t)fn * a.cap;
            for (unsigned r = 0; r < rounds; ++r) {
                const unsigned i = (r * gridDim.x + blockIdx.x) * RUN_BLOCK + threadIdx.x;
                bool to_near = false, keep = false;
                unsigned st = 0;
                if (i < items) {
                    st = __ldcg(far_cur + i);
                    const double lu = __longlong_as_double((long long)__ldcg(&a.st[st].lab));
                    if (lu < thr_old) {
                        // was scanned through the near path with this very label (or is a duplicate of such an entry)
                    } else if (lu < thr) {
                        to_near = true;  // duplicates of the entry are dropped when popped
                    } else {
                        keep = true;
                        const unsigned long long lb = (unsigned long long)__double_as_longlong(lu);
                        fmin = lb < fmin ? lb : fmin;
                    }
                }
                wstage_push(to_near, st, w_near);
                wstage_push(keep, st, w_far);
                if (w_near.n > WSTAGE - 32 * UNROLL) wstage_flush(w_near, near_nxt, &c->ncnt[nxt], a.cap);
                if (w_far.n > WSTAGE - 32 * UNROLL) wstage_flush(w_far, far_nxt, &c->fcnt[fn], a.cap);
            }
            wstage_flush(w_near, near_nxt, &c->ncnt[nxt], a.cap);
            wstage_flush(w_far, far_nxt, &c->fcnt[fn], a.cap);
            far_slot = fn;
            fp = fn;
            advprev = true;
        }
        // one far-minimum update per block and step
        if (__any_sync(0xffffffffu, fmin != ~0ull)) {
            const unsigned long long mn = warp_min_bits(0xffffffffu, fmin);
            if ((threadIdx.x & 31) == 0) atomicMin(&s_fmin, mn);
        }
        __syncthreads();
        if (threadIdx.x == 0 && s_fmin != ~0ull) {
            atomicMin(&c->farmin[far_slot], s_fmin);
            s_fmin = ~0ull;
        }
        grid_barrier(a.bar, gridDim.x);
    }
    for (int o = 16; o; o >>= 1) n_eval += __shfl_xor_sync(0xffffffffu, n_eval, o);
    if ((threadIdx.x & 31) == 0 && n_eval) atomicAdd(a.evaluated, n_eval);
    if (tid == 0) a.status[1] = s;
}

// Predecessors + certificate, one thread per (tree, link) state, writes the tree-major predecessor rows.
struct FinArgs {
    const int *pre_off, *pre;     // in-neighbour links (reverse of succ, legality already applied)
    const int *link_dn;
    const unsigned char *row_ok;  // per link: table row non-negative and non-decreasing
    const double *T;
    int P;
    double periodlength;
    unsigned L, nt;
    const int *rootlink, *rootnode;  // [nt]
    const double *entry;
    const State *st;
    int *pred_out;  // [nt][L]
    int *fail;      // [nt]
};
// grid: x over the links (256 per block), y over the trees (grid-stride) — no 64-bit division per state
__global__ void __launch_bounds__(256) k_fast_finalize(FinArgs a)
{
    const unsigned v = blockIdx.x * blockDim.x + threadIdx.x;
    if (v >= a.L) return;
    // per link, the same for every tree: down node, in-neighbour list, row property
    const int dnv = __ldg(a.link_dn + v);
    const int pb = __ldg(a.pre_off + v), pe = __ldg(a.pre_off + v + 1);
    const bool row_ok = __ldg(a.row_ok + v) != 0;
    for (unsigned t = blockIdx.y; t < a.nt; t += gridDim.y) {
        const size_t st = (size_t)t * a.L + v;
        const unsigned s = (unsigned)__ldg(a.rootlink + t);
        const int rootnode = __ldg(a.rootnode + t);
        const double ent = __ldg(a.entry + t);
        const State *lab = a.st + (size_t)t * a.L;
        bool fail = false;
        int pred = MZ_SP_UNSET;
        const double lv = __longlong_as_double((long long)__ldcs(&lab[v].lab));
        if (v == s) {
            pred = MZ_SP_START;
            const int idx = (int)(ent / a.periodlength);
            if (lv != fast_cost(a.T, a.P, a.L, idx, s)) fail = true;
        }
        if (lv < INF_LABEL) {
            const int idx = (int)((ent + lv) / a.periodlength);
            if (idx < 0 || idx >= a.P) fail = true;  // scanning v looks past the table: not FIFO there
        }
        if (dnv == rootnode) {
            if (v != s && lv < INF_LABEL) fail = true;  // never relaxed by the reference (B/Graph.cpp:385)
        } else {
            int cnt = 0, cand = MZ_SP_UNSET;
            // in-neighbours four at a time: their ids, then their labels, then the costs — each round of loads is issued
            // together instead of one dependent chain per neighbour
            for (int j0 = pb; j0 < pe; j0 += 4) {
                int u[4];
                double lu[4], c[4];
#pragma unroll
                for (int k = 0; k < 4; ++k) u[k] = j0 + k < pe ? __ldg(a.pre + j0 + k) : -1;
#pragma unroll
                for (int k = 0; k < 4; ++k)
                    lu[k] = u[k] >= 0 ? __longlong_as_double((long long)__ldg(&lab[u[k]].lab)) : INF_LABEL;
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    c[k] = 0.0;
                    if (lu[k] < INF_LABEL) c[k] = fast_cost(a.T, a.P, a.L, (int)((ent + lu[k]) / a.periodlength), v);
                }
#pragma unroll
                for (int k = 0; k < 4; ++k)
                    if (lu[k] < INF_LABEL) {
                        const double nl = lu[k] + c[k];
                        if (nl < lv) fail = true;  // not a fix-point (or v unreached although the reference relaxes it)
                        else if (nl == lv) {
                            ++cnt;
                            cand = u[k];
                        }
                        if (!row_ok) fail = true;
                    }
            }
            if (v != s) {
                if (lv < INF_LABEL) {
                    if (cnt != 1) fail = true;  // tie (order decides in the reference) or no producer
                    pred = cand;
                } else if (cnt) {
                    fail = true;
                }
            }
        }
        a.pred_out[st] = pred;
        if (fail) a.fail[t] = 1;
    }
}

// table properties for the certificate and the bucket width: row_ok[l], Σ cost, #cost, any negative
__global__ void k_table_props(const double *__restrict__ T, int P, int L, const int *__restrict__ link_up,
                              unsigned char *__restrict__ row_ok, double *__restrict__ sums)
{
    const int l = blockIdx.x * blockDim.x + threadIdx.x;
    double sum = 0.0, cnt = 0.0, neg = 0.0;
    if (l < L) {
        bool ok = true;
        double prev = 0.0;
        for (int p = 0; p < P; ++p) {
            const double c = T[(size_t)l * P + p];
            if (!(c >= 0.0) || signbit(c)) {
                ok = false;
                neg = 1.0;
            }
            if (p && c < prev) ok = false;
            prev = c;
            if (link_up[l] != MZ_SP_UNSET) {
                sum += c;
                cnt += 1.0;
            }
        }
        row_ok[l] = ok ? 1 : 0;
    }
    for (int o = 16; o; o >>= 1) {
        sum += __shfl_xor_sync(0xffffffffu, sum, o);
        cnt += __shfl_xor_sync(0xffffffffu, cnt, o);
        neg += __shfl_xor_sync(0xffffffffu, neg, o);
    }
    if ((threadIdx.x & 31) == 0) {
        atomicAdd(sums + 0, sum);
        atomicAdd(sums + 1, cnt);
        if (neg > 0.0) atomicAdd(sums + 2, neg);
    }
}

}  // namespace mzfast
