// TEST INFRASTRUCTURE — host emulation of the link-update super-step schedule.
// Compiles the SAME per-event code (mezzo_b200/csrc/sim_core.h) and the SAME host driver (sim_host.h) as the CUDA
// product, with a backend that allocates with malloc and runs the nodes of a super-step one after the other. It exists
// so the CPU test-suite (-m "not gpu") can check the window logic, the tie rules and the multi-rank hand-off without a
// GPU. It is built into tests/emu/libmezzo_emu.so by tests/emu/Makefile; nothing under mezzo_b200/ loads or links it.
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fcntl.h>
#include <sys/mman.h>
#include <sys/stat.h>
#include <unistd.h>

#include <new>
#include <string>
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
const char *last_error() { return g_err; }
}  // namespace mz

using namespace mzsim;

namespace {
// Multi-rank emulation: every rank is a PROCESS; the arenas are POSIX shared-memory segments "<prefix><rank>" mapped by
// all ranks — the CPU stand-in for the NVLink peer mapping of the GPU arenas (cudaIpcOpenMemHandle).
static std::string g_shm_prefix;  // set by emu_sim_create_part around sim_create
struct HostBackend {
    struct Ctx {
        std::string shm_prefix;
        void *peer[MZ_MAXR] = {nullptr};
        size_t peer_bytes = 0;
    };
    static int check_device(int) { return MZ_OK; }
    static int set_device(int) { return MZ_OK; }
    static int ctx_create(Ctx &) { return MZ_OK; }
    static void ctx_destroy(Ctx &) {}
    static int alloc(Ctx &, void **p, size_t bytes)
    {
        *p = std::malloc(bytes);
        return *p ? MZ_OK : MZ_ENOMEM;
    }
    static void free(Ctx &, void *p) { std::free(p); }
    static int h2d(Ctx &, void *dst, const void *src, size_t bytes) { std::memcpy(dst, src, bytes); return MZ_OK; }
    static int h2d_worker(Ctx &, int, void *dst, const void *src, size_t bytes) { std::memcpy(dst, src, bytes); return MZ_OK; }
    static int memset(Ctx &, void *dst, int byte, size_t bytes) { std::memset(dst, byte, bytes); return MZ_OK; }
    static int d2d(Ctx &, void *dst, const void *src, size_t bytes) { std::memcpy(dst, src, bytes); return MZ_OK; }
    static int fill64(Ctx &, void *dst, unsigned long long pattern, size_t n, size_t stride)
    {
        for (size_t i = 0; i < n; ++i) ((unsigned long long *)dst)[i * stride] = pattern;
        return MZ_OK;
    }
    static int sync(Ctx &) { return MZ_OK; }
    static int d2h_sync(Ctx &, void *dst, const void *src, size_t bytes) { std::memcpy(dst, src, bytes); return MZ_OK; }
    static void *map_shm(const std::string &name, size_t bytes, bool create)
    {
        int fd = shm_open(name.c_str(), create ? (O_CREAT | O_RDWR) : O_RDWR, 0600);
        if (fd < 0) return nullptr;
        if (create && ftruncate(fd, (off_t)bytes) != 0) {
            close(fd);
            return nullptr;
        }
        void *p = mmap(nullptr, bytes, PROT_READ | PROT_WRITE, MAP_SHARED, fd, 0);
        close(fd);
        return p == MAP_FAILED ? nullptr : p;
    }
    static int arena_alloc(Ctx &c, void **p, size_t bytes, int rank)
    {
        c.shm_prefix = g_shm_prefix;
        if (c.shm_prefix.empty()) {
            *p = std::calloc(1, bytes);
        } else {
            *p = map_shm(c.shm_prefix + std::to_string(rank), bytes, true);
            c.peer_bytes = bytes;
        }
        return *p ? MZ_OK : MZ_ENOMEM;
    }
    static void arena_free(Ctx &c, void *p, size_t bytes)
    {
        if (c.shm_prefix.empty()) std::free(p);
        else munmap(p, bytes);
    }
    static void detach_peers(SimHost<HostBackend> *s)
    {
        for (int r = 0; r < MZ_MAXR; ++r)
            if (s->ctx.peer[r]) {
                munmap(s->ctx.peer[r], s->ctx.peer_bytes);
                s->ctx.peer[r] = nullptr;
            }
        if (!s->ctx.shm_prefix.empty()) shm_unlink((s->ctx.shm_prefix + std::to_string(s->rank)).c_str());
    }
    // one sweep over this rank's nodes: every node that is a local minimum runs its visit
    static bool sweep(const SimDev &d, std::vector<char> &done, long long *n_done)
    {
        bool ran = false;
        for (int i = 0; i < d.n_my; ++i) {
            if (done[i]) continue;
            double t;
            const int st = async_visit(d, d.my_nodes[i], false, -1.0, -1.0, &t);
            if (st == VISIT_DONE) {
                done[i] = 1;
                ++*n_done;
            }
            if (st != VISIT_WAIT) ran = true;
        }
        return ran;
    }
    // a few regions, so the CPU suite runs through the region tables and sharing classes the kernels use (MZ_EMU_REGIONS overrides)
    static int default_regions(SimHost<HostBackend> *s)
    {
        const char *e = std::getenv("MZ_EMU_REGIONS");
        return e ? std::atoi(e) : std::max(1, std::min(5, (int)s->h_my_nodes.size() / 3));
    }
    static int launch_final(Ctx &, const SimDev &d, int final_node)
    {
        double t;
        async_visit(d, final_node, true, -1.0, -1.0, &t);
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
int emu_sim_create_part(const MzNet *net, int n_ranks, int rank, const int32_t *node_rank, const char *shm_prefix,
                        void **out)
{
    EmuSim *s = nullptr;
    SimPart p;
    p.n_ranks = n_ranks;
    p.rank = rank;
    p.node_rank = node_rank;
    g_shm_prefix = shm_prefix ? shm_prefix : "";
    int rc = sim_create<HostBackend>(0, net, p, &s);
    g_shm_prefix.clear();
    *out = s;
    return rc;
}
int emu_sim_attach_peer(void *h, int peer_rank)
{
    EmuSim *s = (EmuSim *)h;
    MZ_REQUIRE(peer_rank >= 0 && peer_rank < s->n_ranks && peer_rank != s->rank, MZ_EINVAL, "bad peer rank");
    void *p = HostBackend::map_shm(s->ctx.shm_prefix + std::to_string(peer_rank), s->arena_bytes, false);
    MZ_REQUIRE(p, MZ_ESTATE, "cannot map the arena of rank %d", peer_rank);
    s->ctx.peer[peer_rank] = p;
    s->dev.arena[peer_rank] = (char *)p;
    bool all = true;
    for (int r = 0; r < s->n_ranks; ++r) all = all && s->dev.arena[r] != nullptr;
    s->peers_attached = all;
    return MZ_OK;
}
int emu_sim_min_key(void *h, double *t, double *tp, int32_t *sub, int32_t *node)
{
    Key k;
    if (int rc = sim_min_key<HostBackend>((EmuSim *)h, &k)) return rc;
    *t = k.t;
    *tp = k.tp;
    *sub = k.sub;
    *node = k.id == 0x7fffffff ? -1 : k.id;
    return MZ_OK;
}
int emu_sim_final_event(void *h, int32_t node, double end_time) { return sim_final_event<HostBackend>((EmuSim *)h, node, end_time); }
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
int emu_sim_get_linktimes_final(void *s, double *avg) { return sim_get_linktimes_final<HostBackend>((EmuSim *)s, avg); }
int emu_sim_last_stats(void *s, mz_sim_stats *out) { *out = ((EmuSim *)s)->stats; return MZ_OK; }
int emu_sim_run_until(void *s, double t) { return sim_run_until<HostBackend>((EmuSim *)s, t); }
int emu_sim_stop_arrivals(void *s, mz_stop_visit *out, int64_t cap, int64_t *n) { return sim_stop_arrivals<HostBackend>((EmuSim *)s, out, cap, n); }
int emu_sim_stop_departures(void *s, int64_t n, const int32_t *veh, const double *t) { return sim_stop_departures<HostBackend>((EmuSim *)s, n, veh, t); }
int emu_sim_destroy(void *s) { return sim_destroy<HostBackend>((EmuSim *)s); }
}
