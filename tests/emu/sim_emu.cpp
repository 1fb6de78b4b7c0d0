// TEST INFRASTRUCTURE — host emulation of the link-update super-step schedule.
// Compiles the SAME per-event code (mezzo_b200/csrc/sim_core.h) and the SAME host driver (sim_host.h) as the CUDA
// product, with a backend that allocates with malloc and runs the nodes of a super-step one after the other. It exists
// so the CPU test-suite (-m "not gpu") can check the window logic, the tie rules and the multi-rank hand-off without a
// GPU. It is built into tests/emu/libmezzo_emu.so by tests/emu/Makefile; nothing under mezzo_b200/ loads or links it.
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <new>

#include "../../mezzo_b200/csrc/sim_host.h"

namespace mz {
static thread_local char g_err[512] = "";
void set_error(const char *fmt, ...)
{
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}
}  // namespace mz

using namespace mzsim;

namespace {
struct HostBackend {
    struct Ctx {};
    static int check_device(int) { return MZ_OK; }
    static int set_device(int) { return MZ_OK; }
    static int ctx_create(Ctx &) { return MZ_OK; }
    static void ctx_destroy(Ctx &) {}
    static int alloc(void **p, size_t bytes)
    {
        *p = std::malloc(bytes);
        return *p ? MZ_OK : MZ_ENOMEM;
    }
    static void free(void *p) { std::free(p); }
    static int h2d(Ctx &, void *dst, const void *src, size_t bytes) { std::memcpy(dst, src, bytes); return MZ_OK; }
    static int memset(Ctx &, void *dst, int byte, size_t bytes) { std::memset(dst, byte, bytes); return MZ_OK; }
    static int sync(Ctx &) { return MZ_OK; }
    static int d2h_sync(Ctx &, void *dst, const void *src, size_t bytes) { std::memcpy(dst, src, bytes); return MZ_OK; }
    static int arena_alloc(Ctx &, void **p, size_t bytes, int)
    {
        *p = std::calloc(1, bytes);
        return *p ? MZ_OK : MZ_ENOMEM;
    }
    static void arena_free(Ctx &, void *p, size_t) { std::free(p); }
    static void detach_peers(SimHost<HostBackend> *) {}
    static bool window(const SimDev &d, int parity, int final_node)
    {
        // nodes active in one window are pairwise non-adjacent and the published keys are double-buffered,
        // so running them sequentially is equivalent to running them concurrently
        bool alive = false;
        for (int i = 0; i < d.n_my; ++i) alive |= superstep_node(d, d.my_nodes[i], parity, final_node);
        return alive;
    }
    static int launch(Ctx &, const SimDev &d, int parity, int final_node)
    {
        window(d, parity, final_node);
        return MZ_OK;
    }
    static int run_windows(SimHost<HostBackend> *s, long long *steps, long long *launches)
    {
        MZ_REQUIRE(s->n_ranks == 1, MZ_ESTATE, "multi-rank emulation is driven window by window (emu_sim_window)");
        for (;;) {
            const bool alive = window(s->dev, s->parity, -1);
            s->parity ^= 1;
            ++*steps;
            ++*launches;
            if (!alive) break;
        }
        return MZ_OK;
    }
    static int timer_start(Ctx &) { return MZ_OK; }
    static int timer_stop(Ctx &, double *ms) { *ms = 0.0; return MZ_OK; }
};
typedef SimHost<HostBackend> EmuSim;
}  // namespace

extern "C" {
const char *emu_last_error(void) { return mz::g_err; }
int emu_sim_create(const MzNet *net, void **out)
{
    EmuSim *s = nullptr;
    int rc = sim_create<HostBackend>(0, net, SimPart{}, &s);
    *out = s;
    return rc;
}
int emu_sim_run(void *s) { return sim_run<HostBackend>((EmuSim *)s); }
int emu_sim_reset(void *s) { return sim_reset_state<HostBackend>((EmuSim *)s); }
int emu_sim_get_arrivals(void *h, double *a)
{
    EmuSim *s = (EmuSim *)h;
    std::memcpy(a, s->dev.arrival, sizeof(double) * (size_t)s->n_trips);
    return MZ_OK;
}
int emu_sim_get_exit_log(void *s, MzExit *out, int64_t cap, int64_t *n) { return sim_get_exit_log<HostBackend>((EmuSim *)s, out, cap, n); }
int emu_sim_get_linktimes(void *h, double *avg)
{
    EmuSim *s = (EmuSim *)h;
    std::memcpy(avg, s->dev.avgtimes, sizeof(double) * (size_t)s->n_links * s->n_periods);
    return MZ_OK;
}
int emu_sim_last_stats(void *s, mz_sim_stats *out) { *out = ((EmuSim *)s)->stats; return MZ_OK; }
int emu_sim_destroy(void *s) { return sim_destroy<HostBackend>((EmuSim *)s); }
}
