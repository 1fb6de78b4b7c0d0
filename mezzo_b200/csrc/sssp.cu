// Hot path 2: route generation — Graph<double,LinkTimeInfo>::labelCorrecting and friends on sm_100a.
//
// Reference (B/ = /root/reference/branches/busmezzo/mezzo_lib/src/):
//   B/Graph.cpp:313-454  labelCorrecting      B/Graph.cpp:458-483  calcCostToNodes
//   B/Graph.cpp:653-685  shortest_path_vector B/linktimes.cpp:36-121 LinkTime(Info)::cost
//   B/network.cpp:6667-6774 shortest_paths_all (the caller)
//
// Device layout (all in HBM, built once per graph):
//   succ_off[L+1], succ[E] = {v, dnNode(v)}   per-link successor list: the out-links of dnNode(u) in the reference's
//                                             dnLinks_ order, with the static legality test (dnLegal_/dnIndex_, 8-bit
//                                             truncated) already applied — at most 8 entries per link by construction.
//   pre_off[L+1], pre[E]                      reverse of succ (in-neighbour links), ascending u.
//   up_off[N+1], up_list[]                    GraphNode::upLinks_ for calcCostToNodes.
//   T[L][P]                                   dense link-time table, absent periods already replaced by 0.1.
// Per-tree state: label[L] f64, queue[L] i32 (work space), pred[L] i32 (output).
#include <algorithm>
#include <chrono>
#include <cmath>

#include "common.h"
#include "sssp_fast.cuh"

namespace {

constexpr int SP_UNSET = MZ_SP_UNSET;
constexpr int SP_START = MZ_SP_START;
constexpr double SP_INF = MZ_SP_INFINITY;
constexpr int GROUP = 8;  // lanes cooperating on one tree in the exact kernel (= max legal successors per link)

// ---------------------------------------------------------------------------------------------
// device helpers
// ---------------------------------------------------------------------------------------------

// LinkTimeInfo::cost(i,t) on the dense table (B/linktimes.cpp:108-121, 61-82). cvt.rzi saturates where
// the x86 reference yields INT_MIN; both land outside [0,P) and return 0.1.
__device__ __forceinline__ double link_cost(const double *__restrict__ T, int P, double periodlength, int l, double t)
{
    int i = (int)(t / periodlength);
    if (i < 0 || i >= P) return MZ_MISSING_COST;
    return __ldg(T + (size_t)l * P + i);
}

__global__ void k_sanitize_table(double *T, size_t n)
{
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    size_t stride = (size_t)gridDim.x * blockDim.x;
    for (; i < n; i += stride) {
        double v = T[i];
        if (isnan(v)) T[i] = MZ_MISSING_COST;  // "times.count(i) ? times[i] : 0.1"
    }
}

// label=INF, queue=UNSET, pred=UNSET for a chunk of trees (B/Graph.cpp:348-352), flat over chunk*L slots.
__global__ void k_tree_init(double *__restrict__ label, int *__restrict__ queue, int *__restrict__ pred, size_t n)
{
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    size_t stride = (size_t)gridDim.x * blockDim.x;
    // 2 slots per thread per step: 16-byte label stores, 8-byte int stores (buffers are 16B-aligned cudaMalloc rows)
    size_t n2 = n / 2;
    double2 *l2 = reinterpret_cast<double2 *>(label);
    int2 *q2 = reinterpret_cast<int2 *>(queue);
    const bool pred_al = (reinterpret_cast<uintptr_t>(pred) & 7) == 0;
    int2 *p2 = reinterpret_cast<int2 *>(pred);
    for (size_t j = i; j < n2; j += stride) {
        l2[j] = make_double2(SP_INF, SP_INF);
        q2[j] = make_int2(SP_UNSET, SP_UNSET);
        if (pred) {
            if (pred_al) p2[j] = make_int2(SP_UNSET, SP_UNSET);
            else { pred[2 * j] = SP_UNSET; pred[2 * j + 1] = SP_UNSET; }
        }
    }
    if (i == 0 && (n & 1)) {
        label[n - 1] = SP_INF;
        queue[n - 1] = SP_UNSET;
        if (pred) pred[n - 1] = SP_UNSET;
    }
}

// pred rows of the trees listed in slot[] = UNSET (exact re-computation of trees the batched kernel could not certify)
__global__ void k_rows_fill(int *__restrict__ out, const int *__restrict__ slot, int n_rows, int L)
{
    const size_t total = (size_t)n_rows * L;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
        const int r = (int)(i / L);
        out[(size_t)slot[r] * L + (i - (size_t)r * L)] = SP_UNSET;
    }
}

// Exact emulation of the reference's deque label-correcting, one tree per 8-lane group (4 trees per warp).
// The 8 lanes relax the (<= 8) legal successors of the popped link in parallel — they are distinct links, so
// their label/pred updates commute — and then replay the queue operations in successor order so the sequence
// list evolves exactly as in B/Graph.cpp:412-431 (including the literal "queue[rear] = v" overwrite).
__global__ void __launch_bounds__(256) k_sssp_exact(
    const int *__restrict__ succ_off, const int2 *__restrict__ succ, const int *__restrict__ link_up,
    const double *__restrict__ T, int P, double periodlength, int L, int n_trees,
    const int *__restrict__ root, const double *__restrict__ entry,
    double *label_ws, int *queue_ws, int *pred_out, unsigned long long *evaluated, const int *__restrict__ slot)
{
    const int gtid = blockIdx.x * blockDim.x + threadIdx.x;
    const int tree = gtid / GROUP;
    const int lane = threadIdx.x & (GROUP - 1);
    const int gshift = (threadIdx.x & 31) & ~(GROUP - 1);
    const unsigned gmask = ((1u << GROUP) - 1u) << gshift;
    if (tree >= n_trees) return;  // whole groups leave together

    double *label = label_ws + (size_t)tree * L;
    int *queue = queue_ws + (size_t)tree * L;
    int *pred = pred_out + (size_t)(slot ? slot[tree] : tree) * L;  // slot: output row (work space stays compact)

    const int s = root[tree];
    const double ent = entry[tree];
    const int rootnode = __ldg(link_up + s);
    if (lane == 0) {  // B/Graph.cpp:356-363
        pred[s] = SP_START;
        label[s] = link_cost(T, P, periodlength, s, ent);
        queue[s] = L;
    }
    __syncwarp(gmask);

    int u = s, rear = s;
    unsigned n_eval = 0;
    while (u != L) {
        // pop u (B/Graph.cpp:368-369); every lane reads the same words (one broadcast transaction)
        int front = __ldcg(queue + u);
        const double lu = __ldcg(label + u);
        const int b = __ldg(succ_off + u), e = __ldg(succ_off + u + 1);
        __syncwarp(gmask);
        if (lane == 0) queue[u] = SP_START;

        // relax: lane i owns successor i
        int v = -1, qv = 0;
        bool improved = false;
        if (b + lane < e) {
            const int2 sv = __ldg(succ + b + lane);
            v = sv.x;
            if (sv.y != rootnode) {  // B/Graph.cpp:385 — links pointing back into the root's up-node are skipped
                const double nl = lu + link_cost(T, P, periodlength, v, ent + lu);  // B/Graph.cpp:399
                ++n_eval;
                if (nl < __ldcg(label + v)) {  // strict '<', B/Graph.cpp:412
                    improved = true;
                    qv = (v == u) ? SP_START : __ldcg(queue + v);
                    __stcg(label + v, nl);
                    pred[v] = u;
                }
            }
        }
        unsigned bal = (__ballot_sync(gmask, improved) >> gshift) & ((1u << GROUP) - 1u);
        // replay queue operations in successor order (B/Graph.cpp:418-431); all lanes track front/rear
        const int rear0 = rear;
        bool appended = false;
        while (bal) {
            const int j = __ffs(bal) - 1;
            bal &= bal - 1;
            const int vj = __shfl_sync(gmask, v, j, GROUP);
            int qj = __shfl_sync(gmask, qv, j, GROUP);
            // queue[rear0] was overwritten by the first append of this batch: whatever vj==rear0 held before,
            // it now holds a link index ("in list") — this is what the sequential code would read.
            if (appended && vj == rear0) qj = 0;
            if (qj == SP_UNSET) {  // never in list: append at the rear
                if (lane == 0) {
                    __stcg(queue + rear, vj);  // also executed when rear is no longer in the list
                    __stcg(queue + vj, L);
                }
                rear = vj;
                if (front == L) front = vj;
                appended = true;
            } else if (qj == SP_START) {  // was in list before: push to the FRONT
                if (lane == 0) __stcg(queue + vj, front);
                front = vj;
            }
        }
        __syncwarp(gmask);  // order this iteration's label/queue stores before the next pop's loads
        u = front;
    }
    // one atomic per warp-group leader is plenty: trees are long
    if (n_eval) atomicAdd(evaluated, (unsigned long long)n_eval);
}

// calcCostToNodes (B/Graph.cpp:458-483): first minimum over upLinks_ in stored order.
__global__ void k_node_pred(const int *__restrict__ up_off, const int *__restrict__ up_list, int N, int L, int n_trees,
                            const double *__restrict__ label_ws, int *__restrict__ node_pred,
                            double *__restrict__ node_cost, const int *__restrict__ slot, int stride)
{
    size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    size_t total = (size_t)n_trees * N;
    if (idx >= total) return;
    const int tree = (int)(idx / N);
    const int n = (int)(idx - (size_t)tree * N);
    const double *label = label_ws + (size_t)tree * L * stride;  // stride 2: labels inside the frontier kernel's 16-byte states
    double best = SP_INF;
    int p = SP_UNSET;
    const int b = __ldg(up_off + n), e = __ldg(up_off + n + 1);
    for (int i = b; i < e; ++i) {
        const int u = __ldg(up_list + i);
        const double l = __ldcs(label + (size_t)u * stride);
        if (l < best) {
            best = l;
            p = u;
        }
    }
    const size_t o = slot ? (size_t)slot[tree] * N + n : idx;
    if (node_pred) node_pred[o] = p;
    if (node_cost) node_cost[o] = best;
}

// canonical relaxations of a tree = Σ_{reached u} |{v ∈ succ(u) : dnNode(v) != rootnode}|   
__global__ void k_count_canonical(const int *__restrict__ succ_off, const int2 *__restrict__ succ,
                                  const int *__restrict__ link_up, const int *__restrict__ root, int L, int n_trees,
                                  const int *__restrict__ pred, unsigned long long *canonical)
{
    size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    size_t total = (size_t)n_trees * L;
    unsigned c = 0, r = 0;
    if (idx < total) {
        const int tree = (int)(idx / L);
        const int u = (int)(idx - (size_t)tree * L);
        if (__ldcs(pred + idx) != SP_UNSET) {
            r = 1;
            const int rootnode = __ldg(link_up + __ldg(root + tree));
            const int b = __ldg(succ_off + u), e = __ldg(succ_off + u + 1);
            for (int i = b; i < e; ++i) c += (__ldg(succ + i).y != rootnode);
        }
    }
    for (int o = 16; o; o >>= 1) {
        c += __shfl_xor_sync(0xffffffffu, c, o);
        r += __shfl_xor_sync(0xffffffffu, r, o);
    }
    // block-level reduction first: two atomics per block instead of two per warp on the same two addresses
    __shared__ unsigned sc[8], sr[8];
    const int w = threadIdx.x >> 5;
    if ((threadIdx.x & 31) == 0) {
        sc[w] = c;
        sr[w] = r;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        unsigned tc = 0, tr = 0;
        for (int i = 0; i < (int)(blockDim.x >> 5); ++i) {
            tc += sc[i];
            tr += sr[i];
        }
        if (tr) {
            atomicAdd(canonical, (unsigned long long)tc);
            atomicAdd(canonical + 1, (unsigned long long)tr);
        }
    }
}

// reachable + shortest_path_vector + root prepend (B/network.cpp:6726-6742, B/Graph.cpp:653-685), pass 1: lengths.
// A walk longer than L links can only be a cycle ("circular path" → empty route).
__global__ void k_route_len(int n_pairs, const int *__restrict__ pair_tree, const int *__restrict__ dest,
                            const int *__restrict__ root, int L, int N, const int *__restrict__ link_pred,
                            const int *__restrict__ node_pred, int tree0, int n_trees, int *__restrict__ len_out)
{
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_pairs) return;
    const int t = pair_tree[i] - tree0;
    if (t < 0 || t >= n_trees) return;
    const int *lp = link_pred + (size_t)t * L;
    int p = node_pred[(size_t)t * N + dest[i]];
    int len = 0, first = -1;
    while (p >= 0 && len <= L) {
        first = p;
        ++len;
        p = lp[p];
    }
    if (len > L) len = 0;
    else if (len > 0 && first != root[t + tree0]) ++len;
    len_out[i] = len;
}

__global__ void k_route_write(int n_pairs, const int *__restrict__ pair_tree, const int *__restrict__ dest,
                              const int *__restrict__ root, int L, int N, const int *__restrict__ link_pred,
                              const int *__restrict__ node_pred, int tree0, int n_trees,
                              const long long *__restrict__ path_off, long long cap, int *__restrict__ path_links)
{
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_pairs) return;
    const int t = pair_tree[i] - tree0;
    if (t < 0 || t >= n_trees) return;
    const long long b = path_off[i], e = path_off[i + 1];
    if (e == b || e > cap) return;
    const int *lp = link_pred + (size_t)t * L;
    int p = node_pred[(size_t)t * N + dest[i]];
    long long w = e - 1;
    while (p >= 0 && w >= b) {
        path_links[w--] = p;
        p = lp[p];
    }
    if (w == b) path_links[b] = root[t + tree0];  // root link was not the first link: prepend it
}

}  // namespace

// ---------------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------------

struct mz_graph {
    int device = 0;
    int32_t N = 0, L = 0;
    int64_t E = 0;  // successor entries
    std::vector<int32_t> h_up;  // host copy for root validation
    mz::DevBuf<int> succ_off, pre_off, up_off, up_list, link_up, pre;
    mz::DevBuf<int2> succ;
    mz::DevBuf<double> T, Tt;  // link-major (the reference's rows) and period-major (frontier kernels)
    int P = 0;
    double periodlength = 0.0;
    cudaStream_t s_compute = nullptr, s_copy = nullptr;
    cudaStream_t s_own = nullptr;  // the engine-created compute stream (s_compute may point at a caller's stream)
    // per-call device inputs
    mz::DevBuf<int> d_root;
    mz::DevBuf<double> d_entry;
    // double-buffered chunk work space
    struct Chunk {
        mz::DevBuf<double> label, node_cost;
        mz::DevBuf<int> queue, pred, node_pred;
        cudaEvent_t done_compute = nullptr, done_copy = nullptr;
    } chunk[2];
    mz::DevBuf<unsigned long long> d_counters;  // [0]=evaluated [1]=canonical [2]=reached links
    // batched (lanes = trees) kernel: certificate inputs and work space (sssp_fast.cuh)
    mz::DevBuf<int> link_dn;
    mz::DevBuf<unsigned char> row_ok;
    bool uniform_succ = true;   // every in-link of a node has the same legal successor set (no turning prohibitors)
    bool table_nonneg = true;   // no negative cost anywhere in the table
    double mean_cost = 1.0;
    double delta_factor = 8.0;  // bucket width = delta_factor * mean cost (measured best on B200: 3x 39 ms, 8x 33 ms, 16x 31 ms with 10 % extra work)
    unsigned max_steps = 1u << 26;
    size_t list_cap = (size_t)1 << 27;  // entries per near/far list (x6 lists x 4 B); overflow => exact kernel
    struct Fast {
        mz::DevBuf<mzfast::State> st;
        mz::DevBuf<unsigned> nearl, farl, misc;
        mz::DevBuf<mzfast::Ctl> ctl;
        mz::DevBuf<int> rootnode;    // [trees of the chunk]
        mz::DevBuf<int> fail;        // [trees of the chunk]
        mz::DevBuf<double> redo_label, redo_entry;
        mz::DevBuf<int> redo_queue, redo_root, redo_slot;
        int run_blocks = 0;
    } fast;
    unsigned long long fast_steps = 0;  // near-far steps of the last call (diagnostics)
    std::vector<cudaEvent_t> ev;                 // timing events (reused)
    mz_sssp_stats stats{};
    size_t max_chunk_trees = 0;  // 0 = auto
    bool count_canonical = true;
};

namespace {

int graph_build(mz_graph *g, int32_t n_added, const int32_t *add_order, const int32_t *up, const int32_t *dn,
                const int8_t *legal_in)
{
    const int32_t N = g->N, L = g->L;
    std::vector<int32_t> order;
    if (add_order) {
        order.assign(add_order, add_order + n_added);
    } else {
        for (int32_t i = 0; i < L; ++i)
            if (up[i] != SP_UNSET) order.push_back(i);
    }
    std::vector<char> seen(L, 0);
    for (int32_t l : order) {
        MZ_REQUIRE(l >= 0 && l < L, MZ_EINVAL, "add_order holds link id %d outside [0,%d)", l, L);
        MZ_REQUIRE(!seen[l], MZ_EINVAL, "link id %d added twice", l);
        seen[l] = 1;
        MZ_REQUIRE(up[l] >= 0 && up[l] < N && dn[l] >= 0 && dn[l] < N, MZ_EINVAL,
                   "link %d has end nodes (%d,%d) outside [0,%d)", l, up[l], dn[l], N);
    }
    for (int32_t i = 0; i < L; ++i)
        MZ_REQUIRE(seen[i] || up[i] == SP_UNSET, MZ_EINVAL,
                   "link slot %d has end nodes but is not in add_order", i);

    // GraphNode::dnLinks_/upLinks_ in addLink order, dnIndex_ = position in dnLinks_ (B/Graph.cpp:61-62, 689-706)
    std::vector<int32_t> dn_off(N + 1, 0), up_off(N + 1, 0);
    for (int32_t l : order) {
        dn_off[up[l] + 1]++;
        up_off[dn[l] + 1]++;
    }
    for (int32_t i = 0; i < N; ++i) {
        MZ_REQUIRE(dn_off[i + 1] <= 127, MZ_ELIMIT,
                   "node %d has %d out-links; dnIndex_ is a char in the reference (B/Graph.h:103)", i, dn_off[i + 1]);
        dn_off[i + 1] += dn_off[i];
        up_off[i + 1] += up_off[i];
    }
    std::vector<int32_t> dn_list(order.size()), up_list(order.size()), dnidx(L, 0);
    {
        std::vector<int32_t> dc(N, 0), uc(N, 0);
        for (int32_t l : order) {
            dnidx[l] = dc[up[l]];
            dn_list[dn_off[up[l]] + dc[up[l]]++] = l;
            up_list[up_off[dn[l]] + uc[dn[l]]++] = l;
        }
    }
    // successor lists with the static legality filter: (char)(dnLegal_ & (1<<k)) != 0 (B/Graph.h:127-129)
    std::vector<int32_t> succ_off(L + 1, 0);
    std::vector<int2> succ;
    succ.reserve(order.size() * 3);
    for (int32_t u = 0; u < L; ++u) {
        if (seen[u]) {
            const int lg = legal_in ? (int)legal_in[u] : -1;
            const int32_t pivot = dn[u];
            for (int32_t i = dn_off[pivot]; i < dn_off[pivot + 1]; ++i) {
                const int32_t v = dn_list[i];
                const int k = dnidx[v];
                if (k >= 8 || !((lg >> k) & 1)) continue;
                succ.push_back(make_int2(v, dn[v]));
            }
        }
        succ_off[u + 1] = (int32_t)succ.size();
        MZ_REQUIRE(succ_off[u + 1] - succ_off[u] <= GROUP, MZ_ELIMIT, "internal: more than 8 legal successors");
    }
    g->E = (int64_t)succ.size();
    // all in-links of a node share one successor set? (false as soon as a turning prohibitor splits them)
    g->uniform_succ = true;
    {
        std::vector<int32_t> first_in(N, -1);
        for (int32_t u = 0; u < L && g->uniform_succ; ++u) {
            if (!seen[u]) continue;
            int32_t &f = first_in[dn[u]];
            if (f < 0) {
                f = u;
                continue;
            }
            const int32_t n0 = succ_off[f + 1] - succ_off[f], n1 = succ_off[u + 1] - succ_off[u];
            if (n0 != n1) g->uniform_succ = false;
            for (int32_t i = 0; i < n0 && g->uniform_succ; ++i)
                if (succ[succ_off[f] + i].x != succ[succ_off[u] + i].x) g->uniform_succ = false;
        }
    }
    // reverse lists
    std::vector<int32_t> pre_off(L + 1, 0), pre(succ.size());
    for (const int2 &sv : succ) pre_off[sv.x + 1]++;
    for (int32_t i = 0; i < L; ++i) pre_off[i + 1] += pre_off[i];
    {
        std::vector<int32_t> c(L, 0);
        for (int32_t u = 0; u < L; ++u)
            for (int32_t i = succ_off[u]; i < succ_off[u + 1]; ++i) {
                const int32_t v = succ[i].x;
                pre[pre_off[v] + c[v]++] = u;
            }
    }
    g->h_up.assign(up, up + L);
    MZ_CUDA(g->succ_off.upload(succ_off.data(), succ_off.size()));
    MZ_CUDA(g->succ.upload(succ.data(), succ.size()));
    MZ_CUDA(g->pre_off.upload(pre_off.data(), pre_off.size()));
    MZ_CUDA(g->pre.upload(pre.data(), pre.size()));
    MZ_CUDA(g->up_off.upload(up_off.data(), up_off.size()));
    MZ_CUDA(g->up_list.upload(up_list.data(), up_list.size()));
    MZ_CUDA(g->link_up.upload(g->h_up.data(), g->h_up.size()));
    MZ_CUDA(g->link_dn.upload(dn, (size_t)L));
    MZ_CUDA(cudaDeviceSynchronize());
    return MZ_OK;
}

bool use_fast(const mz_graph *g, int mode, size_t n)
{
    return mode == MZ_SSSP_AUTO && g->uniform_succ && g->table_nonneg && n >= 1;
}

size_t pick_chunk_trees(const mz_graph *g, size_t n, bool fast)
{
    // the frontier kernel indexes (tree, link) states with 32 bits
    const size_t hard_cap = fast ? std::max<size_t>(1, (((size_t)1 << 32) - 1) / (size_t)g->L) : (size_t)1 << 30;
    if (g->max_chunk_trees) return std::min(std::min(n, g->max_chunk_trees), hard_cap);
    size_t free_b = 0, total_b = 0;
    if (cudaMemGetInfo(&free_b, &total_b) != cudaSuccess) {
        cudaGetLastError();
        free_b = (size_t)8 << 30;
    }
    if (fast) {
        // per tree: label 8 B + stamp/flag 8 B per link, + outputs (pred 4 B per link, 12 B per node) double-buffered;
        // the six work lists are bounded by list_cap entries each. What the graph already holds from earlier calls — the
        // state array, the lists, the chunk outputs — is reused, so it counts as available (without this a second call on a
        // warm graph saw "free" memory shrunk by its own 40+ GB of state and split a batch that fits into two chunks).
        const double per_tree = (double)g->L * (16.0 + 4.0 * 2) + (double)g->N * 12.0 * 2;
        const double lists = 24.0 * (double)g->list_cap;
        double held = (double)g->fast.st.n * sizeof(mzfast::State) + 4.0 * ((double)g->fast.nearl.n + (double)g->fast.farl.n);
        for (int b = 0; b < 2; ++b)
            held += 4.0 * ((double)g->chunk[b].pred.n + (double)g->chunk[b].node_pred.n) + 8.0 * (double)g->chunk[b].node_cost.n;
        size_t fit = (size_t)std::max(1.0, (((double)free_b + held) * 0.6 - lists) / per_tree);
        return std::min(n, std::min(fit, hard_cap));
    }
    const size_t per_tree = (size_t)g->L * 16 + (size_t)g->N * 12;
    // two chunk buffers, leave 40% of what is free to everybody else
    size_t fit = (size_t)((double)free_b * 0.6 / 2.0 / (double)per_tree);
    // enough trees to fill every SM with resident 8-lane groups (148 SMs × 2048 threads / 8), two waves
    const size_t wave = 148u * 2048u / GROUP;
    fit = std::max<size_t>(1, std::min(fit, 2 * wave));
    return std::min(n, fit);
}

int ensure_events(mz_graph *g, size_t n)
{
    while (g->ev.size() < n) {
        cudaEvent_t e;
        MZ_CUDA(cudaEventCreate(&e));
        g->ev.push_back(e);
    }
    return MZ_OK;
}

struct Outputs {
    int32_t *link_pred;
    int32_t *node_pred;
    double *node_cost;
};

// Exact deque kernel for `cnt` trees. slot == nullptr: trees tree0.. of d_root/d_entry, outputs rows 0..cnt-1 of
// pred/npred/ncost. slot != nullptr: compact root/entry arrays, outputs scattered to rows slot[i].
int launch_exact(mz_graph *g, int cnt, const int *d_root, const double *d_entry, double *label, int *queue, int *pred,
                 int *npred, double *ncost, const int *slot)
{
    const int32_t L = g->L, N = g->N;
    const size_t slots = (size_t)cnt * L;
    k_tree_init<<<(unsigned)std::min<size_t>((slots / 2 + 255) / 256 + 1, 148 * 16), 256, 0, g->s_compute>>>(
        label, queue, slot ? nullptr : pred, slots);
    if (slot)
        k_rows_fill<<<(unsigned)std::min<size_t>((slots + 255) / 256, 148 * 16), 256, 0, g->s_compute>>>(pred, slot, cnt, L);
    const int threads = 256;
    const int blocks = (cnt * GROUP + threads - 1) / threads;
    k_sssp_exact<<<blocks, threads, 0, g->s_compute>>>(g->succ_off.p, g->succ.p, g->link_up.p, g->T.p, g->P,
                                                     g->periodlength, L, cnt, d_root, d_entry, label, queue, pred,
                                                     g->d_counters.p, slot);
    if (npred || ncost) {
        const size_t total = (size_t)cnt * N;
        k_node_pred<<<(unsigned)((total + 255) / 256), 256, 0, g->s_compute>>>(g->up_off.p, g->up_list.p, N, L, cnt, label,
                                                                            npred, ncost, slot, 1);
    }
    g->stats.n_launches += 3 + (slot ? 1 : 0);
    g->stats.main_kernel_launches += 1;
    MZ_CUDA(cudaGetLastError());
    return MZ_OK;
}

// Frontier kernel for the chunk's `nt` trees (tree0.. of d_root/d_entry; host copies of their roots/entries given),
// then the exact kernel for the trees whose certificate failed. Outputs rows 0..nt-1 of pred/npred/ncost.
int run_chunk_fast(mz_graph *g, size_t tree0, int nt, const int32_t *h_root, const double *h_entry, int *pred,
                   int *npred, double *ncost, cudaEvent_t e_main0, cudaEvent_t e_main1)
{
    using namespace mzfast;
    const int32_t L = g->L, N = g->N;
    auto &f = g->fast;
    const size_t states = (size_t)nt * L;
    MZ_REQUIRE(states < ((size_t)1 << 32), MZ_ELIMIT, "internal: chunk of %d trees x %d links exceeds 2^32 states", nt, L);
    std::vector<int> h_rootnode(nt);
    for (int i = 0; i < nt; ++i) h_rootnode[i] = g->h_up[h_root[i]];
    const size_t cap = std::max<size_t>((size_t)nt, std::min<size_t>(states, g->list_cap));
    MZ_CUDA(f.st.ensure(states));
    MZ_CUDA(f.nearl.ensure(3 * cap));
    MZ_CUDA(f.farl.ensure(3 * cap));
    MZ_CUDA(f.ctl.ensure(1));
    MZ_CUDA(f.misc.ensure(8));
    MZ_CUDA(f.fail.ensure(nt));
    MZ_CUDA(f.rootnode.upload(h_rootnode.data(), h_rootnode.size(), g->s_compute));
    MZ_CUDA(cudaMemsetAsync(f.fail.p, 0, sizeof(int) * nt, g->s_compute));

    RunArgs a{};
    a.succ_off = g->succ_off.p;
    a.succ = g->succ.p;
    a.T = g->Tt.p;
    a.P = g->P;
    a.periodlength = g->periodlength;
    a.L = (unsigned)L;
    a.nt = (unsigned)nt;
    a.rootnode = f.rootnode.p;
    a.entry = g->d_entry.p + tree0;
    a.st = f.st.p;
    a.nearl = f.nearl.p;
    a.farl = f.farl.p;
    a.cap = (unsigned)cap;
    a.ctl = f.ctl.p;
    a.bar = f.misc.p;
    a.status = f.misc.p + 2;
    a.evaluated = g->d_counters.p;
    a.delta = g->delta_factor * g->mean_cost;
    a.max_steps = g->max_steps;
    if (f.run_blocks == 0) {
        int per_sm = 0, dev_sms = 0;
        MZ_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_fast_run, RUN_BLOCK, 0));
        MZ_CUDA(cudaDeviceGetAttribute(&dev_sms, cudaDevAttrMultiProcessorCount, g->device));
        MZ_REQUIRE(per_sm > 0, MZ_ECUDA, "k_fast_run does not fit on an SM");
        f.run_blocks = per_sm * dev_sms;
    }
    k_fast_init<<<148 * 8, 256, 0, g->s_compute>>>(f.st.p, states);
    k_fast_ctl_init<<<1, 1, 0, g->s_compute>>>(a);
    k_fast_seed<<<(nt + 255) / 256, 256, 0, g->s_compute>>>(a, g->d_root.p + tree0);
    MZ_CUDA(cudaEventRecord(e_main0, g->s_compute));
    {
        void *params[] = {&a};
        MZ_CUDA(cudaLaunchCooperativeKernel((const void *)k_fast_run, dim3(f.run_blocks), dim3(RUN_BLOCK), params, 0,
                                            g->s_compute));
    }
    MZ_CUDA(cudaEventRecord(e_main1, g->s_compute));
    FinArgs fa{};
    fa.pre_off = g->pre_off.p;
    fa.pre = g->pre.p;
    fa.link_dn = g->link_dn.p;
    fa.row_ok = g->row_ok.p;
    fa.T = g->Tt.p;
    fa.P = g->P;
    fa.periodlength = g->periodlength;
    fa.L = (unsigned)L;
    fa.nt = (unsigned)nt;
    fa.rootlink = g->d_root.p + tree0;
    fa.rootnode = f.rootnode.p;
    fa.entry = g->d_entry.p + tree0;
    fa.st = f.st.p;
    fa.pred_out = pred;
    fa.fail = f.fail.p;
    k_fast_finalize<<<dim3((unsigned)((L + 255) / 256), (unsigned)std::min<size_t>(nt, 65535)), 256, 0, g->s_compute>>>(fa);
    if (npred || ncost) {
        const size_t total = (size_t)nt * N;
        k_node_pred<<<(unsigned)((total + 255) / 256), 256, 0, g->s_compute>>>(g->up_off.p, g->up_list.p, N, L, nt,
                                                                            reinterpret_cast<const double *>(f.st.p),
                                                                            npred, ncost, nullptr, 2);
    }
    g->stats.n_launches += 6;
    g->stats.main_kernel_launches += 1;
    MZ_CUDA(cudaGetLastError());

    // which trees failed their certificate?
    std::vector<int> h_fail(nt);
    unsigned h_status[2] = {0, 0};
    MZ_CUDA(cudaMemcpyAsync(h_fail.data(), f.fail.p, sizeof(int) * nt, cudaMemcpyDeviceToHost, g->s_compute));
    MZ_CUDA(cudaMemcpyAsync(h_status, f.misc.p + 2, sizeof(h_status), cudaMemcpyDeviceToHost, g->s_compute));
    MZ_CUDA(cudaStreamSynchronize(g->s_compute));
    g->fast_steps += h_status[1];
    std::vector<int> redo;
    for (int t = 0; t < nt; ++t)
        if (h_fail[t] || h_status[0]) redo.push_back(t);
    g->stats.n_exact_redo += (int32_t)redo.size();
    if (redo.empty()) return MZ_OK;
    // exact re-computation, scattered into the same output rows
    size_t free_b = 0, total_b = 0;
    if (cudaMemGetInfo(&free_b, &total_b) != cudaSuccess) {
        cudaGetLastError();
        free_b = (size_t)2 << 30;
    }
    size_t R = std::max<size_t>(1, std::min<size_t>(redo.size(), (size_t)((double)free_b * 0.5 / (12.0 * L))));
    R = std::min<size_t>(R, 2 * 148u * 2048u / GROUP);
    if (f.redo_label.n >= (size_t)L) R = std::max(R, std::min(redo.size(), f.redo_label.n / L));
    MZ_CUDA(f.redo_label.ensure(R * L));
    MZ_CUDA(f.redo_queue.ensure(R * L));
    for (size_t r0 = 0; r0 < redo.size(); r0 += R) {
        const int cnt = (int)std::min(R, redo.size() - r0);
        std::vector<int> rr(cnt);
        std::vector<double> re(cnt);
        for (int i = 0; i < cnt; ++i) {
            rr[i] = h_root[redo[r0 + i]];
            re[i] = h_entry[redo[r0 + i]];
        }
        MZ_CUDA(f.redo_root.upload(rr.data(), cnt, g->s_compute));
        MZ_CUDA(f.redo_entry.upload(re.data(), cnt, g->s_compute));
        MZ_CUDA(f.redo_slot.upload(redo.data() + r0, cnt, g->s_compute));
        if (int rc = launch_exact(g, cnt, f.redo_root.p, f.redo_entry.p, f.redo_label.p, f.redo_queue.p, pred, npred,
                                  ncost, f.redo_slot.p))
            return rc;
        MZ_CUDA(cudaStreamSynchronize(g->s_compute));  // the staging vectors above go out of scope
    }
    return MZ_OK;
}

// Runs all trees chunk by chunk; per chunk calls `after(chunk_index, tree0, n_trees, pred, node_pred)` on the compute
// stream (used by the route extraction) and copies requested outputs back on the copy stream while the next chunk computes.
template <class After>
int run_trees(mz_graph *g, int32_t n, const int32_t *root, const double *entry, int mode, Outputs out, After after)
{
    MZ_REQUIRE(g && n >= 0, MZ_EINVAL, "bad graph handle or tree count");
    MZ_REQUIRE(mode == MZ_SSSP_EXACT || mode == MZ_SSSP_AUTO, MZ_EINVAL, "unknown mode %d", mode);
    MZ_REQUIRE(g->P > 0, MZ_ESTATE, "mz_linktimes_set must be called before mz_sssp_*");
    MZ_CUDA(cudaSetDevice(g->device));
    auto t_wall0 = std::chrono::steady_clock::now();
    g->stats = mz_sssp_stats{};
    g->fast_steps = 0;
    if (n == 0) return MZ_OK;
    MZ_REQUIRE(root && entry, MZ_EINVAL, "root/entry must not be null");
    const int32_t L = g->L, N = g->N;

    // host copy of the roots for validation (the reference would index out of bounds)
    std::vector<int32_t> h_root(n);
    std::vector<double> h_entry(n);
    MZ_CUDA(cudaMemcpy(h_root.data(), root, sizeof(int32_t) * n, cudaMemcpyDefault));
    MZ_CUDA(cudaMemcpy(h_entry.data(), entry, sizeof(double) * n, cudaMemcpyDefault));
    for (int32_t i = 0; i < n; ++i)
        MZ_REQUIRE(h_root[i] >= 0 && h_root[i] < L && g->h_up[h_root[i]] != SP_UNSET, MZ_EINVAL,
                   "root[%d]=%d is not a link of the graph", i, h_root[i]);
    MZ_CUDA(g->d_root.upload(h_root.data(), n, g->s_compute));
    MZ_CUDA(g->d_entry.upload(h_entry.data(), n, g->s_compute));
    MZ_CUDA(g->d_counters.ensure(4));
    MZ_CUDA(cudaMemsetAsync(g->d_counters.p, 0, 4 * sizeof(unsigned long long), g->s_compute));

    const bool fast = use_fast(g, mode, (size_t)n);
    const bool lp_dev = mz::is_device_ptr(out.link_pred);
    const bool np_dev = mz::is_device_ptr(out.node_pred);
    const bool nc_dev = mz::is_device_ptr(out.node_cost);
    const size_t C = pick_chunk_trees(g, (size_t)n, fast);
    const size_t n_chunks = ((size_t)n + C - 1) / C;
    const int nbuf = n_chunks > 1 ? 2 : 1;
    for (int b = 0; b < nbuf; ++b) {
        auto &c = g->chunk[b];
        if (!fast) {
            MZ_CUDA(c.label.ensure(C * L));
            MZ_CUDA(c.queue.ensure(C * L));
        }
        if (!lp_dev) MZ_CUDA(c.pred.ensure(C * L));
        if (!np_dev) MZ_CUDA(c.node_pred.ensure(C * N));
        if (!nc_dev && out.node_cost) MZ_CUDA(c.node_cost.ensure(C * N));
    }
    if (int rc = ensure_events(g, 4 * n_chunks)) return rc;

    for (size_t ci = 0; ci < n_chunks; ++ci) {
        auto &c = g->chunk[ci & 1];
        const size_t tree0 = ci * C;
        const int nt = (int)std::min(C, (size_t)n - tree0);
        // buffers of this slot are free once their previous copy-out finished
        if (ci >= 2) MZ_CUDA(cudaStreamWaitEvent(g->s_compute, c.done_copy, 0));
        int *pred = lp_dev ? out.link_pred + tree0 * L : c.pred.p;
        int *npred = np_dev ? out.node_pred + tree0 * N : c.node_pred.p;
        double *ncost = nc_dev ? out.node_cost + tree0 * N : (out.node_cost ? c.node_cost.p : nullptr);
        cudaEvent_t e0 = g->ev[4 * ci], e1 = g->ev[4 * ci + 1], e2 = g->ev[4 * ci + 2], e3 = g->ev[4 * ci + 3];

        MZ_CUDA(cudaEventRecord(e0, g->s_compute));
        if (fast) {
            if (int rc = run_chunk_fast(g, tree0, nt, h_root.data() + tree0, h_entry.data() + tree0, pred, npred, ncost, e1, e2))
                return rc;
        } else {
            MZ_CUDA(cudaEventRecord(e1, g->s_compute));
            if (int rc = launch_exact(g, nt, g->d_root.p + tree0, g->d_entry.p + tree0, c.label.p, c.queue.p, pred, npred,
                                      ncost, nullptr))
                return rc;
            MZ_CUDA(cudaEventRecord(e2, g->s_compute));
        }
        if (g->count_canonical) {
            const size_t slots = (size_t)nt * L;
            k_count_canonical<<<(unsigned)((slots + 255) / 256), 256, 0, g->s_compute>>>(
                g->succ_off.p, g->succ.p, g->link_up.p, g->d_root.p + tree0, L, nt, pred, g->d_counters.p + 1);
            g->stats.n_launches++;
        }
        if (int rc = after(ci, tree0, nt, pred, npred)) return rc;
        MZ_CUDA(cudaEventRecord(e3, g->s_compute));
        MZ_CUDA(cudaGetLastError());
        MZ_CUDA(cudaEventRecord(c.done_compute, g->s_compute));

        // copy-out on the second stream overlaps the next chunk's kernels
        MZ_CUDA(cudaStreamWaitEvent(g->s_copy, c.done_compute, 0));
        const size_t slots = (size_t)nt * L;
        if (out.link_pred && !lp_dev)
            MZ_CUDA(cudaMemcpyAsync(out.link_pred + tree0 * L, pred, sizeof(int32_t) * slots, cudaMemcpyDeviceToHost,
                                    g->s_copy));
        if (out.node_pred && !np_dev)
            MZ_CUDA(cudaMemcpyAsync(out.node_pred + tree0 * N, npred, sizeof(int32_t) * (size_t)nt * N,
                                    cudaMemcpyDeviceToHost, g->s_copy));
        if (out.node_cost && !nc_dev)
            MZ_CUDA(cudaMemcpyAsync(out.node_cost + tree0 * N, ncost, sizeof(double) * (size_t)nt * N,
                                    cudaMemcpyDeviceToHost, g->s_copy));
        MZ_CUDA(cudaEventRecord(c.done_copy, g->s_copy));
    }
    MZ_CUDA(cudaStreamSynchronize(g->s_compute));
    MZ_CUDA(cudaStreamSynchronize(g->s_copy));

    unsigned long long h_cnt[4] = {0, 0, 0, 0};
    MZ_CUDA(cudaMemcpy(h_cnt, g->d_counters.p, sizeof(h_cnt), cudaMemcpyDeviceToHost));
    g->stats.evaluated_relax = (int64_t)h_cnt[0];
    g->stats.canonical_relax = g->count_canonical ? (int64_t)h_cnt[1] : -1;
    g->stats.reached_links = g->count_canonical ? (int64_t)h_cnt[2] : -1;
    for (size_t ci = 0; ci < n_chunks; ++ci) {
        float ms_all = 0, ms_main = 0;
        MZ_CUDA(cudaEventElapsedTime(&ms_all, g->ev[4 * ci], g->ev[4 * ci + 3]));
        MZ_CUDA(cudaEventElapsedTime(&ms_main, g->ev[4 * ci + 1], g->ev[4 * ci + 2]));
        g->stats.kernel_ms += ms_all;
        g->stats.main_kernel_ms += ms_main;
    }
    g->stats.n_chunks = (int32_t)n_chunks;
    g->stats.fast_steps = (int64_t)g->fast_steps;
    g->stats.total_ms =
        std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t_wall0).count();
    return MZ_OK;
}

}  // namespace

extern "C" {

int mz_graph_create(int device, int32_t n_node_slots, int32_t n_link_slots, int32_t n_added,
                    const int32_t *add_order, const int32_t *up_node, const int32_t *dn_node,
                    const int8_t *dn_legal, mz_graph **out)
{
    MZ_REQUIRE(out, MZ_EINVAL, "out handle pointer is null");
    *out = nullptr;
    MZ_REQUIRE(n_node_slots > 0 && n_link_slots > 0 && up_node && dn_node, MZ_EINVAL,
               "node/link slot counts must be positive and up_node/dn_node non-null");
    MZ_REQUIRE(!add_order || n_added >= 0, MZ_EINVAL, "n_added negative");
    int ndev = 0;
    {
        cudaError_t e = cudaGetDeviceCount(&ndev);
        if (e != cudaSuccess || ndev == 0) {
            cudaGetLastError();
            mz::set_error("no CUDA device available (%s); mezzo_b200 has no CPU fallback",
                          e == cudaSuccess ? "device count 0" : cudaGetErrorString(e));
            return MZ_ENODEV;
        }
    }
    MZ_REQUIRE(device >= 0 && device < ndev, MZ_EINVAL, "device %d outside [0,%d)", device, ndev);
    MZ_CUDA(cudaSetDevice(device));
    mz_graph *g = new (std::nothrow) mz_graph;
    MZ_REQUIRE(g, MZ_ENOMEM, "out of host memory");
    g->device = device;
    g->N = n_node_slots;
    g->L = n_link_slots;
    int rc = graph_build(g, n_added, add_order, up_node, dn_node, dn_legal);
    if (rc == MZ_OK) {
        cudaError_t e = cudaStreamCreateWithFlags(&g->s_own, cudaStreamNonBlocking);
        g->s_compute = g->s_own;
        if (e == cudaSuccess) e = cudaStreamCreateWithFlags(&g->s_copy, cudaStreamNonBlocking);
        for (int b = 0; b < 2 && e == cudaSuccess; ++b) {
            e = cudaEventCreateWithFlags(&g->chunk[b].done_compute, cudaEventDisableTiming);
            if (e == cudaSuccess) e = cudaEventCreateWithFlags(&g->chunk[b].done_copy, cudaEventDisableTiming);
        }
        if (e != cudaSuccess) rc = mz::cuda_fail(e, "stream/event creation", __FILE__, __LINE__);
    }
    if (rc != MZ_OK) {
        mz_graph_destroy(g);
        return rc;
    }
    *out = g;
    return MZ_OK;
}

int mz_linktimes_set(mz_graph *g, int32_t n_periods, double periodlength, const double *T)
{
    MZ_REQUIRE(g && T, MZ_EINVAL, "null graph or table");
    MZ_REQUIRE(n_periods > 0 && periodlength > 0.0, MZ_EINVAL, "n_periods and periodlength must be positive");
    MZ_CUDA(cudaSetDevice(g->device));
    const size_t n = (size_t)g->L * n_periods;
    MZ_CUDA(g->T.upload(T, n, g->s_compute));
    k_sanitize_table<<<(unsigned)std::min<size_t>((n + 255) / 256, 148 * 8), 256, 0, g->s_compute>>>(g->T.p, n);
    MZ_CUDA(cudaGetLastError());
    MZ_CUDA(g->Tt.ensure(n));
    mzfast::k_transpose_table<<<(unsigned)std::min<size_t>((n + 255) / 256, 148 * 8), 256, 0, g->s_compute>>>(g->T.p, g->Tt.p, n_periods,
                                                                                                          (unsigned)g->L);
    MZ_CUDA(cudaGetLastError());
    // certificate inputs of the batched kernel: FIFO rows, sign of the costs, mean cost (bucket width)
    {
        mz::DevBuf<double> sums;
        MZ_CUDA(sums.ensure(4));
        MZ_CUDA(cudaMemsetAsync(sums.p, 0, 4 * sizeof(double), g->s_compute));
        MZ_CUDA(g->row_ok.ensure((size_t)g->L));
        mzfast::k_table_props<<<(g->L + 255) / 256, 256, 0, g->s_compute>>>(g->T.p, n_periods, g->L, g->link_up.p,
                                                                            g->row_ok.p, sums.p);
        MZ_CUDA(cudaGetLastError());
        double h[4] = {0, 0, 0, 0};
        MZ_CUDA(cudaMemcpyAsync(h, sums.p, sizeof(h), cudaMemcpyDeviceToHost, g->s_compute));
        MZ_CUDA(cudaStreamSynchronize(g->s_compute));
        g->table_nonneg = h[2] == 0.0;
        g->mean_cost = (h[1] > 0.0 && h[0] > 0.0 && std::isfinite(h[0])) ? h[0] / h[1] : 1.0;
    }
    g->P = n_periods;
    g->periodlength = periodlength;
    return MZ_OK;
}

int mz_sssp_batch(mz_graph *g, int32_t n, const int32_t *root, const double *entry, int mode, int32_t *link_pred,
                  int32_t *node_pred, double *node_cost)
{
    Outputs o{link_pred, node_pred, node_cost};
    return run_trees(g, n, root, entry, mode, o, [](size_t, size_t, int, int *, int *) { return MZ_OK; });
}

int mz_sssp_routes(mz_graph *g, int32_t n, const int32_t *root, const double *entry, int mode,
                   const int64_t *dest_off, const int32_t *dest, int64_t *path_off, int32_t *path_links,
                   int64_t path_cap, int64_t *n_path_links)
{
    MZ_REQUIRE(g && dest_off && path_off && n_path_links && n >= 0, MZ_EINVAL, "null argument");
    MZ_CUDA(cudaSetDevice(g->device));
    const int64_t n_pairs64 = n ? dest_off[n] : 0;
    MZ_REQUIRE(n_pairs64 >= 0 && n_pairs64 < (int64_t)INT32_MAX, MZ_ELIMIT, "too many (tree,dest) pairs");
    const int n_pairs = (int)n_pairs64;
    *n_path_links = 0;
    path_off[0] = 0;
    if (n_pairs == 0) {
        Outputs o{nullptr, nullptr, nullptr};
        return run_trees(g, n, root, entry, mode, o, [](size_t, size_t, int, int *, int *) { return MZ_OK; });
    }
    MZ_REQUIRE(dest && (path_links || path_cap == 0), MZ_EINVAL, "null dest/path_links");
    std::vector<int32_t> h_pair_tree(n_pairs);
    for (int32_t t = 0; t < n; ++t) {
        MZ_REQUIRE(dest_off[t] <= dest_off[t + 1], MZ_EINVAL, "dest_off not monotone");
        for (int64_t i = dest_off[t]; i < dest_off[t + 1]; ++i) {
            MZ_REQUIRE(dest[i] >= 0 && dest[i] < g->N, MZ_EINVAL, "dest node %d outside [0,%d)", dest[i], g->N);
            h_pair_tree[i] = t;
        }
    }
    mz::DevBuf<int> d_pair_tree, d_dest, d_len, d_paths;
    mz::DevBuf<long long> d_off;
    MZ_CUDA(d_pair_tree.upload(h_pair_tree.data(), n_pairs, g->s_compute));
    MZ_CUDA(d_dest.upload(dest, n_pairs, g->s_compute));
    MZ_CUDA(d_len.ensure(n_pairs));
    MZ_CUDA(cudaMemsetAsync(d_len.p, 0, sizeof(int) * n_pairs, g->s_compute));

    // Pass 1 over all chunks: path lengths. The trees are recomputed for pass 2 only when they do not fit one chunk.
    std::vector<int> h_len(n_pairs);
    std::vector<long long> h_off(n_pairs + 1, 0);
    const int L = g->L, N = g->N;
    auto lens = [&](size_t, size_t tree0, int nt, int *pred, int *npred) {
        k_route_len<<<(n_pairs + 255) / 256, 256, 0, g->s_compute>>>(n_pairs, d_pair_tree.p, d_dest.p, g->d_root.p, L,
                                                                    N, pred, npred, (int)tree0, nt, d_len.p);
        g->stats.n_launches++;
        return MZ_OK;
    };
    Outputs o{nullptr, nullptr, nullptr};
    if (int rc = run_trees(g, n, root, entry, mode, o, lens)) return rc;
    mz_sssp_stats st1 = g->stats;
    // decided from what run_trees DID (its chunk size comes from the free memory at that moment): with one chunk the
    // predecessor arrays of all n trees are still resident in chunk[0]
    const bool single = st1.n_chunks == 1;
    MZ_CUDA(cudaMemcpy(h_len.data(), d_len.p, sizeof(int) * n_pairs, cudaMemcpyDeviceToHost));
    for (int i = 0; i < n_pairs; ++i) h_off[i + 1] = h_off[i] + h_len[i];
    for (int i = 0; i <= n_pairs; ++i) path_off[i] = h_off[i];
    *n_path_links = h_off[n_pairs];
    MZ_REQUIRE(h_off[n_pairs] <= path_cap, MZ_ELIMIT, "path buffer too small: need %lld links, cap %lld",
               (long long)h_off[n_pairs], (long long)path_cap);
    if (h_off[n_pairs] == 0) return MZ_OK;
    MZ_CUDA(d_off.upload(h_off.data(), n_pairs + 1, g->s_compute));
    const bool paths_dev = mz::is_device_ptr(path_links);
    int *d_out = path_links;
    if (!paths_dev) {
        MZ_CUDA(d_paths.ensure((size_t)h_off[n_pairs]));
        d_out = d_paths.p;
    }
    auto writes = [&](size_t, size_t tree0, int nt, int *pred, int *npred) {
        k_route_write<<<(n_pairs + 255) / 256, 256, 0, g->s_compute>>>(n_pairs, d_pair_tree.p, d_dest.p, g->d_root.p, L,
                                                                      N, pred, npred, (int)tree0, nt, d_off.p,
                                                                      (long long)path_cap, d_out);
        g->stats.n_launches++;
        return MZ_OK;
    };
    if (single) {
        // the single chunk's pred arrays are still resident in chunk[0]
        if (int rc = writes(0, 0, n, g->chunk[0].pred.p, g->chunk[0].node_pred.p)) return rc;
        MZ_CUDA(cudaGetLastError());
        MZ_CUDA(cudaStreamSynchronize(g->s_compute));
        g->stats = st1;
        g->stats.n_launches++;
    } else {
        if (int rc = run_trees(g, n, root, entry, mode, o, writes)) return rc;
        g->stats.kernel_ms += st1.kernel_ms;
        g->stats.main_kernel_ms += st1.main_kernel_ms;
        g->stats.total_ms += st1.total_ms;
        g->stats.n_launches += st1.n_launches;
        g->stats.main_kernel_launches += st1.main_kernel_launches;
    }
    if (!paths_dev)
        MZ_CUDA(cudaMemcpy(path_links, d_out, sizeof(int32_t) * (size_t)h_off[n_pairs], cudaMemcpyDeviceToHost));
    return MZ_OK;
}

int mz_sssp_last_stats(const mz_graph *g, mz_sssp_stats *out)
{
    MZ_REQUIRE(g && out, MZ_EINVAL, "null argument");
    *out = g->stats;
    return MZ_OK;
}

int mz_graph_set_option(mz_graph *g, int option, int64_t value)
{
    MZ_REQUIRE(g, MZ_EINVAL, "null graph");
    switch (option) {
        case MZ_OPT_CHUNK_TREES:
            MZ_REQUIRE(value >= 0, MZ_EINVAL, "chunk size must be >= 0");
            g->max_chunk_trees = (size_t)value;
            return MZ_OK;
        case MZ_OPT_COUNT_CANONICAL:
            g->count_canonical = value != 0;
            return MZ_OK;
        case MZ_OPT_DELTA_MILLI:
            MZ_REQUIRE(value >= 1, MZ_EINVAL, "bucket width factor must be >= 1 (thousandths of the mean link cost)");
            g->delta_factor = (double)value / 1000.0;
            return MZ_OK;
        case MZ_OPT_MAX_STEPS:
            MZ_REQUIRE(value >= 1, MZ_EINVAL, "step cap must be >= 1");
            g->max_steps = (unsigned)std::min<int64_t>(value, 0x7fffffff);
            return MZ_OK;
        default:
            mz::set_error("unknown option %d", option);
            return MZ_EINVAL;
    }
}

int mz_graph_set_stream(mz_graph *g, void *cuda_stream)
{
    MZ_REQUIRE(g, MZ_EINVAL, "null graph");
    g->s_compute = cuda_stream ? (cudaStream_t)cuda_stream : g->s_own;
    return MZ_OK;
}

int mz_graph_destroy(mz_graph *g)
{
    if (!g) return MZ_OK;
    cudaSetDevice(g->device);
    for (auto e : g->ev) cudaEventDestroy(e);
    for (int b = 0; b < 2; ++b) {
        if (g->chunk[b].done_compute) cudaEventDestroy(g->chunk[b].done_compute);
        if (g->chunk[b].done_copy) cudaEventDestroy(g->chunk[b].done_copy);
    }
    if (g->s_own) cudaStreamDestroy(g->s_own);
    if (g->s_copy) cudaStreamDestroy(g->s_copy);
    delete g;
    cudaGetLastError();
    return MZ_OK;
}
}
