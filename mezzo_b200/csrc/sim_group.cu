// Hot path 1, thread-per-node variant of the persistent node-visit kernel (sim.cu holds the warp-per-node one).
//
// The per-event code of sim_core.h is written against a "warp" abstraction whose lane count is a compile-time constant;
// here it is compiled with MZ_GROUP_W = 1 lane, so every THREAD owns nodes and runs whole visits by itself and a hardware
// warp works on up to 32 nodes at the same time. Why: the warp-per-node kernel is latency-bound with 12 resident warps per SM (ncu on a 254 k-link network: 80 % of
// the issue slots have no eligible warp, IPC 0.76) and its queue scans rarely find more than 2-3 vehicles to spread over
// 32 lanes; on networks with far more nodes than resident warps what counts is the number of visits in flight.
// Results are identical: a visit is the same sequential code for any lane count, and the conservative rule (a node runs only
// while its key precedes every neighbour's published key) does not care which lanes of which warp own a node.
// No group ever waits inside a branch for another group: a group checks, visits or moves on, and every warp-level
// primitive names only the group's own lanes — diverged groups of one warp therefore cannot dead-lock each other.
#ifndef MZ_GROUP_W
#define MZ_GROUP_W 1
#endif
// Lane groups of 2-16 lanes (several nodes per warp, queue scans still lane-parallel) compile from the same source, and a
// kernel with ONE active group per warp is bit-exact — but with several groups of a warp active at once the runs fault
// (illegal instruction / illegal address on B200, CUDA 12.9) although the collectives themselves behave in isolation
// (tools/scratch/gtest.cu). Until that is understood only the single-lane build is allowed into the library.
#if MZ_GROUP_W != 1 && !defined(MZ_GROUP_EXPERIMENTAL)
#error "MZ_GROUP_W > 1 is experimental (see the comment above): add -DMZ_GROUP_EXPERIMENTAL to build it"
#endif
#define MZ_SIM_NS mzsimg
#include <algorithm>

#include "common.h"
#include "sim_core.h"

using namespace mzsimg;

#ifndef MZ_GTHREADS
#define MZ_GTHREADS 512  // 128 registers per thread, 16 warps per SM
#endif

namespace {

constexpr int kW = MZ_GROUP_W;
constexpr int kGroupsPerBlock = MZ_GTHREADS / kW;

struct alignas(16) NodeSlotG {  // per owned node: cached key time, node id (-1 = none), done flag
    double keyt;
    int node;
    int done;
};

template <bool SIG>
__global__ void __launch_bounds__(MZ_GTHREADS, 1) k_sim_async_g(const __grid_constant__ SimDev d, unsigned long long max_idle_sweeps, int rounds)
{
    // per-group state of the owned nodes, `rounds` x W slots per group
    extern __shared__ NodeSlotG s_slots_g[];
    const unsigned lane = threadIdx.x & (kW - 1), grp = threadIdx.x / kW;
    // The block owns one region (sim_host.h: sim_set_regions): my_nodes[r0, r1). Slot i = round * groups + group of the block
    // maps to region position (i % 32) * (slots / 32) + i / 32: the lanes of one warp take nodes that lie slots/32 apart in
    // the region's node order. Consecutive nodes are neighbours on the road network, and neighbours can never run at the
    // same time (conservative rule) — consecutive nodes in one warp would leave most lanes waiting for each other; nodes
    // far apart become candidates together and run in SIMT lockstep.
    const unsigned r0 = (unsigned)lds(d.region_off + blockIdx.x), r1 = (unsigned)lds(d.region_off + blockIdx.x + 1);
    const unsigned n_reg = r1 - r0;
    const unsigned slots_blk = (unsigned)rounds * kGroupsPerBlock;  // a multiple of 32
    const unsigned S = (unsigned)rounds * kW;
#if MZ_GROUP_W == 1 && !defined(MZ_GROUP_CONSECUTIVE)
#define MZ_NODE_POS(j) ((((j) * kGroupsPerBlock + grp) & 31u) * (slots_blk / 32u) + (((j) * kGroupsPerBlock + grp) >> 5))
#endif
    // slot (round r, lane l) of this group sits at s_slots_g[r * blockDim.x + threadIdx.x]: conflict-free for any W
    NodeSlotG *slots = s_slots_g + threadIdx.x;
    unsigned n_own = 0;
    for (unsigned j = lane; j < S; j += kW) {
#if MZ_GROUP_W == 1 && !defined(MZ_GROUP_CONSECUTIVE)
        const unsigned pos = MZ_NODE_POS(j / kW);
#else
        const unsigned pos = (j / kW) * MZ_GTHREADS + threadIdx.x;  // round r: the group's W lanes hold W consecutive nodes
#endif
        NodeSlotG ns;
        ns.node = pos < n_reg ? lds(d.my_nodes + r0 + pos) : -1;
        ns.keyt = ns.node >= 0 ? node_key_time(d, ns.node, 2) : INFINITY;
        ns.done = ns.node < 0;
        slots[(j / kW) * MZ_GTHREADS] = ns;
        if (ns.node >= 0) ++n_own;
    }
    mz_syncwarp();
#if MZ_GROUP_W > 1
    // nodes owned by the whole group (every lane leaves the visit loop below at the same moment)
    for (int m = kW / 2; m > 0; m >>= 1) n_own += (unsigned)mz_shfl_xor((int)n_own, m);
#endif
    const unsigned n_rounds = (unsigned)rounds;
    unsigned long long acc[5] = {0, 0, 0, 0, 0};
    unsigned n_done = 0;
    unsigned long long idle = 0;
#ifdef MZ_GROUP_SERIAL  // debugging aid: the groups of a warp take turns, one visit loop at a time
    while (__any_sync(0xffffffffu, n_done < n_own)) {
#else
    while (n_done < n_own) {
#endif
        bool ran = false;
#pragma unroll 1
        for (unsigned r = 0; r < n_rounds; ++r) {
            // lane j pre-checks node j of the round: its cached key time against the neighbours' published times
            const unsigned slot = r * MZ_GTHREADS;
            const NodeSlotG mine = slots[slot];
            const double kt = mine.keyt;
            bool cand = false;
            double nbt = INFINITY;
            if (!mine.done) {
                cand = true;  // (a node past the horizon is a candidate too: the visit retires it)
                if (kt < d.runtime) {
                    const int n = mine.node;
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
                }
            }
            unsigned cm = mz_ballot(cand);
#ifdef MZ_GROUP_SERIAL
            for (unsigned turn = 0; turn < 32u / kW; ++turn) {
                __syncwarp();
                if (((threadIdx.x & 31u) / kW) != turn) continue;
#endif
#pragma unroll 1
            while (cm) {
                const int j = __ffs((int)cm) - 1;
                cm &= cm - 1;
                const int nn = mz_shfl(mine.node, j);
                const double my_t = mz_shfl(kt, j);
                const double nb_bound = mz_shfl(nbt, j);
                double nt;
                const int st = async_visit<SIG>(d, nn, false, my_t, nb_bound, &nt, acc);
                if ((int)lane == j) {
                    slots[slot].keyt = nt;
                    if (st == VISIT_DONE) slots[slot].done = 1;
                }
                if (st == VISIT_DONE) ++n_done;
                if (st != VISIT_WAIT) ran = true;
            }
#ifdef MZ_GROUP_SERIAL
            }
            __syncwarp();
#endif
            mz_syncwarp();
        }
#ifdef MZ_GROUP_SERIAL
        ran = __any_sync(0xffffffffu, ran);
#endif
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
    if (lane == 0)
        for (int i = 0; i < 5; ++i)
            if (acc[i]) atomicAdd(d.counters + i, acc[i]);
}

}  // namespace

// Launch interface for sim.cu (same SimDev layout in both namespaces: it is the same header).
namespace mzgroup {

static const int kSmemMax = 200 * 1024;

int threads_per_block() { return kGroupsPerBlock; }

// resident blocks of the kernel at its largest slot table: a cooperative launch of that many blocks always fits
int max_blocks(int device, int *blocks_out)
{
    int sms = 0, per_sm = 0;
    MZ_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, device));
    MZ_CUDA(cudaFuncSetAttribute(k_sim_async_g<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemMax));
    MZ_CUDA(cudaFuncSetAttribute(k_sim_async_g<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemMax));
    MZ_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_sim_async_g<true>, MZ_GTHREADS, kSmemMax));
    MZ_REQUIRE(per_sm > 0, MZ_ECUDA, "k_sim_async_g does not fit on an SM");
    *blocks_out = per_sm * sms;
    return MZ_OK;
}

// one block per region (d.n_regions of them, at most max_blocks), `max_region` = nodes of the largest region
int launch(const void *simdev, size_t simdev_bytes, int max_region, unsigned long long max_idle, cudaStream_t stream)
{
    MZ_REQUIRE(simdev_bytes == sizeof(SimDev), MZ_EINVAL, "SimDev layout mismatch between the two kernel variants");
    SimDev d;
    std::memcpy(&d, simdev, sizeof(d));
    const void *kern = (d.sig_ev || d.trip_stop_off) ? (const void *)k_sim_async_g<true> : (const void *)k_sim_async_g<false>;
    int rounds = std::max(1, (max_region + MZ_GTHREADS - 1) / MZ_GTHREADS);  // every thread holds one node slot per round
    const size_t smem = (size_t)rounds * MZ_GTHREADS * sizeof(NodeSlotG);
    MZ_REQUIRE(smem <= (size_t)kSmemMax, MZ_ELIMIT, "a region of %d nodes exceeds what one block's lane groups track; use more ranks", max_region);
    void *params[] = {&d, &max_idle, &rounds};
    // cooperative launch only to guarantee that every group is resident (the nodes wait for each other)
    MZ_CUDA(cudaLaunchCooperativeKernel(kern, dim3(d.n_regions), dim3(MZ_GTHREADS), params, smem, stream));
    return MZ_OK;
}

int lanes_per_node() { return kW; }

}  // namespace mzgroup
