"""Host-side mirror of the link-update seam on top of the C ABI (mz_sim_*).

The reference call site is `while (time<runtime) time=eventlist->next();` inside Network::step
(B/network.cpp:7804-7807, B/ = /root/reference/branches/busmezzo/mezzo_lib/src/) over the object graph built by
executemaster + init; here the object graph is the flattened MzNet (mezzo_b200/flatten.py) and the loop is
mz_sim_run on the GPU. Results come back in the reference's terms: OD rows (ODpair::report, B/od.cpp:379-384),
per-vehicle link exit times, per-link period averages (Link::avgtimes)."""
import ctypes as C

import numpy as np

from . import _lib
from ._lib import SimStats, check, lib

EXIT_DT = np.dtype([("vid", "<i4"), ("link", "<i4"), ("entry", "<f8"), ("exit", "<f8")])
VISIT_DT = np.dtype([("veh", "<i4"), ("stop", "<i4"), ("link", "<i4"), ("pad", "<i4"), ("t_enter", "<f8"), ("t_arrive", "<f8")])


def run_with_stops(sim, depart_time, slice_s):
    """The co-simulation loop a transit host runs around the engine: slices of `slice_s` seconds —
    no longer than the host model's shortest dwell time, so that no departure falls into a slice already run — after each
    of which the stop visits are handed to `depart_time(visit) -> time the bus leaves the stop` (the host's
    Busstop::execute, B/busline.cpp:1277-1298) and the answers are booked. Works on anything with run_until /
    stop_arrivals / stop_departures / step (Simulation, and the test-suite's host emulation). Returns all visits."""
    horizon = float(sim.flat.struct.runtime)
    waiting, seen = [], []  # visits whose arrival lies beyond the slice just run stay with the host until their slice
    t = 0.0
    while t < horizon:
        t = min(horizon, t + slice_s)
        sim.run_until(t)
        for v in sim.stop_arrivals():
            waiting.append(v)
            seen.append(v)
        # an arrival inside the slice just run: its bus leaves at least `slice_s` later, i.e. not before t
        now = [v for v in waiting if v["t_arrive"] < t]
        waiting = [v for v in waiting if not v["t_arrive"] < t]
        if now:
            sim.stop_departures([int(v["veh"]) for v in now], [float(depart_time(v)) for v in now])
    sim.step()
    seen.extend(sim.stop_arrivals())
    return np.array(seen, VISIT_DT) if seen else np.empty(0, VISIT_DT)


def _bind():
    l = lib()
    if getattr(l, "_sim_bound", False):
        return l
    vp = C.c_void_p
    l.mz_sim_create.argtypes = [C.c_int, C.POINTER(_lib.MzNetStruct), C.POINTER(vp)]
    l.mz_sim_run.argtypes = [vp]
    l.mz_sim_reset.argtypes = [vp]
    l.mz_sim_get_arrivals.argtypes = [vp, vp]
    l.mz_sim_get_exit_log.argtypes = [vp, vp, C.c_int64, C.POINTER(C.c_int64)]
    l.mz_sim_get_linktimes.argtypes = [vp, vp]
    l.mz_sim_get_linktimes_final.argtypes = [vp, vp]
    l.mz_sim_last_stats.argtypes = [vp, C.POINTER(SimStats)]
    l.mz_sim_set_stream.argtypes = [vp, vp]
    l.mz_sim_set_option.argtypes = [vp, C.c_int, C.c_int64]
    l.mz_sim_destroy.argtypes = [vp]
    l.mz_sim_create_part.argtypes = [C.c_int, C.POINTER(_lib.MzNetStruct), C.c_int, C.c_int, vp, C.POINTER(vp)]
    l.mz_sim_arena_handle.argtypes = [vp, vp, C.POINTER(C.c_int64)]
    l.mz_sim_attach_peer.argtypes = [vp, C.c_int, vp]
    l.mz_sim_min_key.argtypes = [vp, C.POINTER(C.c_double), C.POINTER(C.c_double), C.POINTER(C.c_int32),
                                 C.POINTER(C.c_int32)]
    l.mz_sim_final_event.argtypes = [vp, C.c_int32, C.c_double]
    l.mz_sim_run_until.argtypes = [vp, C.c_double]
    l.mz_sim_stop_arrivals.argtypes = [vp, vp, C.c_int64, C.POINTER(C.c_int64)]
    l.mz_sim_stop_departures.argtypes = [vp, C.c_int64, vp, vp]
    l._sim_bound = True
    return l


def od_rows(f, arr):
    """The rows Vehicle::report appends to ODpair::grid (B/vehicle.cpp:79-91): o, d, vid, start, end, end-start,
    meters, route id, switched — for arrived vehicles, grouped by OD pair (sorted) then arrival order.
    f = FlatNet, arr = arrival time per trip (mz_sim_get_arrivals; < 0 = not arrived)."""
    done = np.nonzero(arr >= 0)[0]
    routes = f.trip_route[done]
    cum = np.concatenate([[0], np.cumsum(f.link_length[f.route_links], dtype=np.int64)])
    meters = (cum[f.route_off[1:]] - cum[f.route_off[:-1]]).astype(np.float64)  # Σ link lengths: integers, order-free
    od = f.trip_od_sorted[done]
    od_o = np.array([x[0] for x in f.od_sorted], np.float64)
    od_d = np.array([x[1] for x in f.od_sorted], np.float64)
    rows = np.stack([od_o[od], od_d[od], done + 1.0, f.trip_time[done], arr[done], arr[done] - f.trip_time[done],
                     meters[routes], f.route_ids[routes].astype(np.float64), np.zeros(len(done))], axis=1)
    order = np.lexsort((arr[done], od))
    return rows[order]


class Simulation:
    """Network::step replacement: Simulation(flat).step() runs the whole horizon on the GPU."""

    def __init__(self, flat, device=0):
        self.flat = flat
        shared = getattr(flat, "shared_stochastic_servers", None)
        if shared:
            some = ", ".join(f"{sid} (nodes {', '.join(map(str, ns[:4]))}{' ...' if len(ns) > 4 else ''})" for sid, ns in list(shared.items())[:5])
            raise _lib.MzError(_lib.MZ_ELIMIT, f"{len(shared)} stochastic server(s) are shared by several nodes: {some}. Their random "
                               "streams follow the global event order (SURVEY.md H5); give each node its own server or use "
                               "deterministic servers")
        self.l = _bind()
        h = C.c_void_p()
        check(self.l.mz_sim_create(device, C.byref(flat.struct), C.byref(h)))
        self._h = h

    def step(self):
        check(self.l.mz_sim_run(self._h))
        return self.stats()["end_time"]

    def reset(self):
        check(self.l.mz_sim_reset(self._h))

    # transit overlay hooks: mz_sim_run_until / mz_sim_stop_arrivals / mz_sim_stop_departures
    def run_until(self, t):
        check(self.l.mz_sim_run_until(self._h, float(t)))

    def stop_arrivals(self):
        n = C.c_int64(0)
        check(self.l.mz_sim_stop_arrivals(self._h, None, 0, C.byref(n)))
        out = np.empty(max(1, n.value), VISIT_DT)
        if n.value:
            check(self.l.mz_sim_stop_arrivals(self._h, out.ctypes.data_as(C.c_void_p), n.value, C.byref(n)))
        return out[:n.value]

    def stop_departures(self, veh, t_depart):
        v = np.ascontiguousarray(veh, np.int32)
        t = np.ascontiguousarray(t_depart, np.float64)
        check(self.l.mz_sim_stop_departures(self._h, len(v), v.ctypes.data_as(C.c_void_p), t.ctypes.data_as(C.c_void_p)))

    def set_stream(self, cuda_stream):
        check(self.l.mz_sim_set_stream(self._h, C.c_void_p(cuda_stream) if cuda_stream else None))

    def set_option(self, option, value):
        check(self.l.mz_sim_set_option(self._h, int(option), int(value)))

    def stats(self):
        st = SimStats()
        check(self.l.mz_sim_last_stats(self._h, C.byref(st)))
        return st.asdict()

    def arrivals(self):
        a = np.empty(max(1, self.flat.struct.n_trips), np.float64)
        check(self.l.mz_sim_get_arrivals(self._h, a.ctypes.data_as(C.c_void_p)))
        return a[:self.flat.struct.n_trips]

    def exit_log(self):
        n = C.c_int64(0)
        check(self.l.mz_sim_get_exit_log(self._h, None, 0, C.byref(n)))
        out = np.empty(max(1, n.value), EXIT_DT)
        check(self.l.mz_sim_get_exit_log(self._h, out.ctypes.data_as(C.c_void_p), n.value, C.byref(n)))
        return out[:n.value]

    def linktimes(self, final=False):
        """Link::avgtimes [L, P]; final=True applies Link::end_of_simulation first (what Link::write_time streams)."""
        L, P = self.flat.struct.n_links, self.flat.struct.n_periods
        a = np.empty((L, P), np.float64)
        fn = self.l.mz_sim_get_linktimes_final if final else self.l.mz_sim_get_linktimes
        check(fn(self._h, a.ctypes.data_as(C.c_void_p)))
        return a

    def od_rows(self):
        return od_rows(self.flat, self.arrivals())

    def close(self):
        if self._h is not None:
            self.l.mz_sim_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class CudaPartBackend:
    """Per-rank engine of mezzo_b200.dist.PartitionedSimulation on a GPU: mz_sim_create_part + the peer-arena calls."""

    def __init__(self, device=0):
        self.device = int(device)
        self.l = _bind()
        self._h = None
        self.flat = None

    def create(self, flat, world, rank, node_rank):
        self.flat = flat
        h = C.c_void_p()
        self._node_rank = np.ascontiguousarray(node_rank, np.int32)
        check(self.l.mz_sim_create_part(self.device, C.byref(flat.struct), int(world), int(rank),
                                        self._node_rank.ctypes.data_as(C.c_void_p), C.byref(h)))
        self._h = h

    def export_handle(self):
        buf = (C.c_char * 64)()
        n = C.c_int64(0)
        check(self.l.mz_sim_arena_handle(self._h, buf, C.byref(n)))
        return bytes(buf)

    def attach(self, peer_rank, handle):
        buf = (C.c_char * 64).from_buffer_copy(handle)
        check(self.l.mz_sim_attach_peer(self._h, int(peer_rank), buf))

    def reset(self):
        check(self.l.mz_sim_reset(self._h))

    def run(self):
        check(self.l.mz_sim_run(self._h))

    def min_key(self):
        t, tp, sub, node = C.c_double(), C.c_double(), C.c_int32(), C.c_int32()
        check(self.l.mz_sim_min_key(self._h, C.byref(t), C.byref(tp), C.byref(sub), C.byref(node)))
        return (t.value, tp.value, sub.value, node.value)

    def final_event(self, node, end_time):
        check(self.l.mz_sim_final_event(self._h, int(node), float(end_time)))

    def set_option(self, option, value):
        check(self.l.mz_sim_set_option(self._h, int(option), int(value)))

    stats = Simulation.stats
    arrivals = Simulation.arrivals
    exit_log = Simulation.exit_log
    linktimes = Simulation.linktimes

    def destroy(self):
        if self._h is not None:
            self.l.mz_sim_destroy(self._h)
            self._h = None
