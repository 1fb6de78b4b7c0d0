// Hot path 1: the link update on sm_100a — CUDA backend + C ABI (mz_sim_*).
// The data layout and per-event functions live in sim_core.h, the backend-independent host logic in sim_host.h.
#include "common.h"
#include "sim_host.h"

using namespace mzsim;

namespace {

// One conservative window: one WARP per node (logical process); queue scans/shifts are spread over the lanes.
// `final_node` >= 0 selects the clean-up launch that executes exactly one event (the first one at/after the
// horizon).
constexpr int kSimThreads = 128;
__global__ void __launch_bounds__(kSimThreads) k_sim_superstep(SimDev d, int parity, int final_node)
{
    const int n = (int)((blockIdx.x * (unsigned)blockDim.x + threadIdx.x) >> 5);
    if (n >= d.n_nodes) return;  // warp-uniform
    superstep_node(d, n, parity, final_node);
}

struct CudaBackend {
    struct Ctx {
        cudaStream_t own = nullptr, stream = nullptr;
        cudaEvent_t ev0 = nullptr, ev1 = nullptr;
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
        c.stream = c.own;
        return MZ_OK;
    }
    static void ctx_destroy(Ctx &c)
    {
        if (c.ev0) cudaEventDestroy(c.ev0);
        if (c.ev1) cudaEventDestroy(c.ev1);
        if (c.own) cudaStreamDestroy(c.own);
        cudaGetLastError();
    }
    static int alloc(void **p, size_t bytes)
    {
        MZ_CUDA(cudaMalloc(p, bytes));
        return MZ_OK;
    }
    static void free(void *p) { cudaFree(p); }
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
        k_sim_superstep<<<(d.n_nodes + nodes_per_block - 1) / nodes_per_block, kSimThreads, 0, c.stream>>>(d, parity,
                                                                                                         final_node);
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
    int rc = sim_create<CudaBackend>(device, net, &s);
    if (out) *out = static_cast<mz_sim *>(s);
    return rc;
}

int mz_sim_reset(mz_sim *s)
{
    MZ_REQUIRE(s, MZ_EINVAL, "null handle");
    return sim_reset_state<CudaBackend>(s);
}

int mz_sim_run(mz_sim *s) { return sim_run<CudaBackend>(s); }

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
        case MZ_SIM_OPT_CHECK_EVERY:
            MZ_REQUIRE(value >= 1 && value <= 4096, MZ_EINVAL, "check interval must be in [1,4096]");
            s->check_every = (int)value;
            return MZ_OK;
        default:
            mz::set_error("unknown option %d", option);
            return MZ_EINVAL;
    }
}

int mz_sim_destroy(mz_sim *s) { return sim_destroy<CudaBackend>(s); }
}
