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
#include <vector>

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
    // one sweep over this rank's nodes: every node that is a local minimum runs its visit
    static bool sweep(const SimDev &d, std::vector<char> &done, long long *n_done)
    {
        bool ran = false;
        for (int i = 0; i < d.n_my; ++i) {
            if (done[i]) continue;
            double t;
            const int st = async_visit(d, d.my_nodes[i], false, &t);
            if (st == VISIT_DONE) {
                done[i] = 1;
                ++*n_done;
            }
            if (st != VISIT_WAIT) ran = true;
        }
        return ran;
    }
    static int launch_final(Ctx &, const SimDev &d, int final_node)
    {
        double t;
        async_visit(d, final_node, true, &t);
        return MZ_OK;
    }
    static int run_visits(SimHost<HostBackend> *s, long long *steps, long long *launches)
    {
        // single process: the sweeps visit the nodes one after the other; several ranks (processes sharing the arenas)
        // sweep concurrently and a rank simply keeps polling while its nodes wait for a neighbour on another rank
        std::vector<char> done((size_t)s->dev.n_my, 0);
        long long n_done = 0, idle = 0;
        while (n_done < s->dev.n_my) {
            const bool ran = sweep(s->dev, done, &n_done);
            ++*steps;
            idle = ran ? 0 : idle + 1;
            MZ_REQUIRE(s->n_ranks > 1 || idle < 2, MZ_ESTATE, "emulation: no node can run (dead-lock in the visit rule)");
            MZ_REQUIRE(idle < (1ll << 34), MZ_ESTATE, "emulation: a peer rank never advanced");
        }
        ++*launches;
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
