"""ctypes bindings for the TEST INFRASTRUCTURE under oracle/ (C restatement + compiled reference).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs import this.
"""
import ctypes as C
import os
import subprocess

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ORACLE_DIR = os.path.join(ROOT, "oracle")
SP_UNSET, SP_START = -2, -1
INF = 9999999.0

_i32p = np.ctypeslib.ndpointer(np.int32, flags="C_CONTIGUOUS")
_i8p = np.ctypeslib.ndpointer(np.int8, flags="C_CONTIGUOUS")
_f64p = np.ctypeslib.ndpointer(np.float64, flags="C_CONTIGUOUS")


def _opt(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


def build_oracle():
    subprocess.check_call(["make", "-s", "-C", ORACLE_DIR, "oracle"])


_orc = None


def oracle_lib():
    global _orc
    if _orc is None:
        path = os.path.join(ORACLE_DIR, "liboracle.so")
        if not os.path.exists(path):
            build_oracle()
        lib = C.CDLL(path)
        lib.orc_graph_create.restype = C.c_void_p
        lib.orc_graph_create.argtypes = [C.c_int32, C.c_int32, C.c_int32, C.c_void_p, _i32p, _i32p, C.c_void_p]
        lib.orc_graph_destroy.argtypes = [C.c_void_p]
        lib.orc_linktimes_set.argtypes = [C.c_void_p, C.c_int32, C.c_double, _f64p]
        lib.orc_label_correcting.restype = C.c_int64
        lib.orc_label_correcting.argtypes = [C.c_void_p, C.c_int32, C.c_double, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
        lib.orc_route.restype = C.c_int32
        lib.orc_route.argtypes = [C.c_void_p, C.c_int32, _i32p, _i32p, C.c_int32, _i32p, C.c_int32]
        lib.orc_sssp_batch.restype = C.c_int64
        lib.orc_sssp_batch.argtypes = [C.c_void_p, C.c_int32, _i32p, _f64p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
        _orc = lib
    return _orc


def ref_sssp_path():
    return os.path.join(ORACLE_DIR, "_ref", "libref_sssp.so")


_ref = None


def ref_lib():
    """The compiled UNMODIFIED reference (oracle/_ref/libref_sssp.so) or None if it was never built."""
    global _ref
    if _ref is None:
        p = ref_sssp_path()
        if not os.path.exists(p):
            return None
        lib = C.CDLL(p)
        lib.ref_graph_create.restype = C.c_void_p
        lib.ref_graph_create.argtypes = [C.c_int, C.c_int]
        lib.ref_graph_add_link.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_double]
        lib.ref_graph_set_downlink_indices.argtypes = [C.c_void_p]
        lib.ref_graph_set_turning_prohibitor.argtypes = [C.c_void_p, C.c_int, C.c_int]
        lib.ref_graph_dn_legal.argtypes = [C.c_void_p, C.c_int]
        lib.ref_linktimes_set.argtypes = [C.c_void_p, C.c_int, C.c_double, _f64p, C.c_void_p]
        lib.ref_label_correcting.restype = C.c_longlong
        lib.ref_label_correcting.argtypes = [C.c_void_p, C.c_int, C.c_double, C.c_void_p, C.c_void_p, C.c_void_p]
        lib.ref_reachable.argtypes = [C.c_void_p, C.c_int]
        lib.ref_shortest_path_vector.argtypes = [C.c_void_p, C.c_int, _i32p, C.c_int]
        lib.ref_graph_destroy.argtypes = [C.c_void_p]
        _ref = lib
    return _ref


class GraphSpec:
    """Plain description of a Graph<double,LinkTimeInfo> build: what Network::init_shortest_path does
    (B/network.cpp:6576-6630) expressed as arrays. Slots never added keep up=dn=SP_UNSET."""

    def __init__(self, n_nodes, n_links, up, dn, add_order=None, prohibitors=()):
        self.n_nodes = int(n_nodes)
        self.n_links = int(n_links)
        self.up = np.ascontiguousarray(up, np.int32)
        self.dn = np.ascontiguousarray(dn, np.int32)
        if add_order is None:
            add_order = np.nonzero(self.up != SP_UNSET)[0]
        self.add_order = np.ascontiguousarray(add_order, np.int32)
        self.prohibitors = [(int(a), int(b)) for a, b in prohibitors]
        self.legal = self._legal_mask()

    def dn_index(self):
        """dnIndex_ per link = position among its up-node's out-links in add order, stored in a char."""
        cnt = {}
        idx = np.zeros(self.n_links, np.int64)
        for l in self.add_order:
            u = int(self.up[l])
            idx[l] = cnt.get(u, 0)
            cnt[u] = idx[l] + 1
        return idx

    def _legal_mask(self):
        """dnLegal_ after set_turning_prohibitor XORs (B/Graph.cpp:709-717): char arithmetic, so
        1<<index with index>=8 leaves the byte unchanged and a repeated prohibitor toggles the bit back."""
        legal = np.full(self.n_links, 0xFF, np.int64)
        idx = self.dn_index()
        for a, b in self.prohibitors:
            legal[a] = (legal[a] ^ (1 << int(idx[b]))) & 0xFF
        return legal.astype(np.uint8).view(np.int8)


class OracleGraph:
    def __init__(self, spec: GraphSpec):
        self.spec = spec
        self.lib = oracle_lib()
        self.h = self.lib.orc_graph_create(spec.n_nodes, spec.n_links, len(spec.add_order), _opt(spec.add_order),
                                           spec.up, spec.dn, _opt(spec.legal))
        self.T = None

    def set_linktimes(self, P, periodlength, T):
        self.T = np.ascontiguousarray(T, np.float64).reshape(self.spec.n_links, P)
        self.lib.orc_linktimes_set(self.h, P, periodlength, self.T)

    def tree(self, root, entry):
        s = self.spec
        lp = np.empty(s.n_links, np.int32)
        npd = np.empty(s.n_nodes, np.int32)
        nc = np.empty(s.n_nodes, np.float64)
        canon = C.c_int64(0)
        ncost = self.lib.orc_label_correcting(self.h, root, entry, _opt(lp), _opt(npd), _opt(nc), C.byref(canon))
        return lp, npd, nc, int(ncost), int(canon.value)

    def batch(self, roots, entries, want_pred=True):
        s = self.spec
        roots = np.ascontiguousarray(roots, np.int32)
        entries = np.ascontiguousarray(entries, np.float64)
        n = len(roots)
        lp = np.empty((n, s.n_links), np.int32) if want_pred else None
        npd = np.empty((n, s.n_nodes), np.int32) if want_pred else None
        nc = np.empty((n, s.n_nodes), np.float64) if want_pred else None
        canon = C.c_int64(0)
        ncost = self.lib.orc_sssp_batch(self.h, n, roots, entries, _opt(lp), _opt(npd), _opt(nc), C.byref(canon))
        return lp, npd, nc, int(ncost), int(canon.value)

    def route(self, root, lp, npd, dest, cap=None):
        cap = cap or (self.spec.n_links + 1)
        out = np.empty(cap, np.int32)
        n = self.lib.orc_route(self.h, root, lp, npd, dest, out, cap)
        assert n >= 0
        return out[:n].copy()

    def __del__(self):
        try:
            self.lib.orc_graph_destroy(self.h)
        except Exception:
            pass


class RefGraph:
    """Drives the compiled reference exactly like Network::init_shortest_path does."""

    def __init__(self, spec: GraphSpec):
        self.spec = spec
        self.lib = ref_lib()
        assert self.lib is not None, "oracle/_ref/libref_sssp.so missing (make -C oracle ref)"
        self.h = self.lib.ref_graph_create(spec.n_nodes, spec.n_links)
        for l in spec.add_order:
            self.lib.ref_graph_add_link(self.h, int(l), int(spec.up[l]), int(spec.dn[l]), 1.0)
        self.lib.ref_graph_set_downlink_indices(self.h)
        for a, b in spec.prohibitors:
            self.lib.ref_graph_set_turning_prohibitor(self.h, a, b)

    def dn_legal(self):
        return np.array([self.lib.ref_graph_dn_legal(self.h, l) for l in range(self.spec.n_links)], np.int8)

    def set_linktimes(self, P, periodlength, T, has_row=None):
        T = np.ascontiguousarray(T, np.float64).reshape(self.spec.n_links, P)
        hr = None if has_row is None else np.ascontiguousarray(has_row, np.int8)
        self.lib.ref_linktimes_set(self.h, P, periodlength, T, _opt(hr))

    def tree(self, root, entry):
        s = self.spec
        lp = np.empty(s.n_links, np.int32)
        npd = np.empty(s.n_nodes, np.int32)
        nc = np.empty(s.n_nodes, np.float64)
        ncost = self.lib.ref_label_correcting(self.h, root, entry, _opt(lp), _opt(npd), _opt(nc))
        return lp, npd, nc, int(ncost)

    def route(self, root, dest):
        """reachable + get_path + root prepend (B/network.cpp:6726-6742) on the last tree."""
        if not self.lib.ref_reachable(self.h, dest):
            return np.empty(0, np.int32)
        out = np.empty(self.spec.n_links + 1, np.int32)
        n = self.lib.ref_shortest_path_vector(self.h, dest, out, len(out))
        assert n >= 0
        path = list(out[:n])
        if n > 0 and path[0] != root:
            path.insert(0, root)
        return np.array(path, np.int32)

    def __del__(self):
        try:
            self.lib.ref_graph_destroy(self.h)
        except Exception:
            pass


# ---------------------------------------------------------------------------------------------------------
# hot path 1: C restatement of the link-update event loop (oracle/sim_oracle.c)
# ---------------------------------------------------------------------------------------------------------
EXIT_DT = np.dtype([("vid", "<i4"), ("link", "<i4"), ("entry", "<f8"), ("exit", "<f8")])


def oracle_sim(flat, dwell=None):
    """Runs the sequential CPU restatement on a FlatNet; returns dict(exits, arrivals, linktimes, stats, seconds).
    dwell (transit hooks, row f3): seconds a bus spends at each stop visit, [n_stops] — the stand-in for the host's transit
    model; the stop visits come back as `visits`."""
    import time
    from mezzo_b200._lib import MzNetStruct, SimStats
    from mezzo_b200.sim import VISIT_DT
    lib = oracle_lib()
    lib.orc_sim_create.restype = C.c_void_p
    lib.orc_sim_create.argtypes = [C.POINTER(MzNetStruct)]
    lib.orc_sim_run.argtypes = [C.c_void_p]
    lib.orc_sim_destroy.argtypes = [C.c_void_p]
    lib.orc_sim_stats.restype = C.POINTER(SimStats)
    lib.orc_sim_stats.argtypes = [C.c_void_p]
    lib.orc_sim_arrivals.restype = C.POINTER(C.c_double)
    lib.orc_sim_arrivals.argtypes = [C.c_void_p]
    lib.orc_sim_linktimes.restype = C.POINTER(C.c_double)
    lib.orc_sim_linktimes.argtypes = [C.c_void_p]
    lib.orc_sim_exit_log.restype = C.c_int64
    lib.orc_sim_exit_log.argtypes = [C.c_void_p, C.POINTER(C.c_void_p)]
    lib.orc_sim_set_dwell.argtypes = [C.c_void_p, C.c_void_p]
    lib.orc_sim_stop_visits.restype = C.c_int64
    lib.orc_sim_stop_visits.argtypes = [C.c_void_p, C.POINTER(C.c_void_p)]
    h = lib.orc_sim_create(C.byref(flat.struct))
    if dwell is not None:
        dwell = np.ascontiguousarray(dwell, np.float64)
        lib.orc_sim_set_dwell(h, dwell.ctypes.data_as(C.c_void_p))
    t0 = time.perf_counter()
    lib.orc_sim_run(h)
    secs = time.perf_counter() - t0
    st = lib.orc_sim_stats(h).contents.asdict()
    nt, L, P = flat.struct.n_trips, flat.struct.n_links, flat.struct.n_periods
    arr = np.ctypeslib.as_array(lib.orc_sim_arrivals(h), (max(nt, 1),))[:nt].copy()
    lt = np.ctypeslib.as_array(lib.orc_sim_linktimes(h), (max(L * P, 1),))[:L * P].reshape(L, P).copy()
    p = C.c_void_p()
    n = lib.orc_sim_exit_log(h, C.byref(p))
    if n:
        buf = (C.c_char * (n * EXIT_DT.itemsize)).from_address(p.value)
        exits = np.frombuffer(buf, EXIT_DT).copy()
    else:
        exits = np.empty(0, EXIT_DT)
    visits = np.empty(0, VISIT_DT)
    if getattr(flat, "trip_stop_off", None) is not None:
        nv = lib.orc_sim_stop_visits(h, C.byref(p))
        if nv:
            visits = np.frombuffer((C.c_char * (nv * VISIT_DT.itemsize)).from_address(p.value), VISIT_DT).copy()
    lib.orc_sim_destroy(h)
    return dict(exits=exits, arrivals=arr, linktimes=lt, stats=st, seconds=secs, visits=visits)


# ---------------------------------------------------------------------------------------------------------
# host emulation of the product's super-step schedule (tests/emu/sim_emu.cpp) — CPU tests only
# ---------------------------------------------------------------------------------------------------------
_emu = None


def emu_lib():
    global _emu
    if _emu is None:
        d = os.path.join(ROOT, "tests", "emu")
        subprocess.check_call(["make", "-s", "-C", d])
        from mezzo_b200._lib import MzNetStruct, SimStats
        lib = C.CDLL(os.path.join(d, "libmezzo_emu.so"))
        lib.emu_last_error.restype = C.c_char_p
        lib.emu_sim_create.argtypes = [C.POINTER(MzNetStruct), C.POINTER(C.c_void_p)]
        for f in ("emu_sim_run", "emu_sim_reset", "emu_sim_destroy"):
            getattr(lib, f).argtypes = [C.c_void_p]
        lib.emu_sim_get_arrivals.argtypes = [C.c_void_p, C.c_void_p]
        lib.emu_sim_get_linktimes.argtypes = [C.c_void_p, C.c_void_p]
        lib.emu_sim_get_linktimes_final.argtypes = [C.c_void_p, C.c_void_p]
        lib.emu_sim_get_exit_log.argtypes = [C.c_void_p, C.c_void_p, C.c_int64, C.POINTER(C.c_int64)]
        lib.emu_sim_last_stats.argtypes = [C.c_void_p, C.POINTER(SimStats)]
        _emu = lib
    return _emu


def emu_sim(flat):
    """Runs the product's per-event code and super-step driver on the CPU (sequential nodes per super-step)."""
    from mezzo_b200._lib import SimStats
    lib = emu_lib()
    h = C.c_void_p()
    rc = lib.emu_sim_create(C.byref(flat.struct), C.byref(h))
    if rc != 0:
        raise RuntimeError(f"emu_sim_create failed {rc}: {lib.emu_last_error().decode()}")
    rc = lib.emu_sim_run(h)
    assert rc == 0, lib.emu_last_error()
    st = SimStats()
    lib.emu_sim_last_stats(h, C.byref(st))
    nt, L, P = flat.struct.n_trips, flat.struct.n_links, flat.struct.n_periods
    arr = np.empty(max(1, nt))
    lib.emu_sim_get_arrivals(h, arr.ctypes.data_as(C.c_void_p))
    lt = np.empty((L, P))
    lib.emu_sim_get_linktimes(h, lt.ctypes.data_as(C.c_void_p))
    ltf = np.empty((L, P))
    lib.emu_sim_get_linktimes_final(h, ltf.ctypes.data_as(C.c_void_p))
    n = C.c_int64(0)
    lib.emu_sim_get_exit_log(h, None, 0, C.byref(n))
    ex = np.empty(max(1, n.value), EXIT_DT)
    lib.emu_sim_get_exit_log(h, ex.ctypes.data_as(C.c_void_p), n.value, C.byref(n))
    lib.emu_sim_destroy(h)
    return dict(exits=ex[:n.value], arrivals=arr[:nt], linktimes=lt, linktimes_final=ltf, stats=st.asdict())


class EmuSimulation:
    """The host emulation behind the interface of mezzo_b200.Simulation (run_until / stop_arrivals / stop_departures /
    step / outputs), so that mezzo_b200.sim.run_with_stops drives it like the GPU engine. Test infrastructure only."""

    def __init__(self, flat):
        from mezzo_b200.sim import VISIT_DT
        self.flat, self.lib, self.VISIT_DT = flat, emu_lib(), VISIT_DT
        self.lib.emu_sim_run_until.argtypes = [C.c_void_p, C.c_double]
        self.lib.emu_sim_stop_arrivals.argtypes = [C.c_void_p, C.c_void_p, C.c_int64, C.POINTER(C.c_int64)]
        self.lib.emu_sim_stop_departures.argtypes = [C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p]
        self.h = C.c_void_p()
        self._ck(self.lib.emu_sim_create(C.byref(flat.struct), C.byref(self.h)))

    def _ck(self, rc):
        if rc != 0:
            raise RuntimeError(f"emu error {rc}: {self.lib.emu_last_error().decode()}")

    def run_until(self, t):
        self._ck(self.lib.emu_sim_run_until(self.h, float(t)))

    def stop_arrivals(self):
        n = C.c_int64(0)
        self._ck(self.lib.emu_sim_stop_arrivals(self.h, None, 0, C.byref(n)))
        out = np.empty(max(1, n.value), self.VISIT_DT)
        if n.value:
            self._ck(self.lib.emu_sim_stop_arrivals(self.h, out.ctypes.data_as(C.c_void_p), n.value, C.byref(n)))
        return out[:n.value]

    def stop_departures(self, veh, t_depart):
        v, t = np.ascontiguousarray(veh, np.int32), np.ascontiguousarray(t_depart, np.float64)
        self._ck(self.lib.emu_sim_stop_departures(self.h, len(v), v.ctypes.data_as(C.c_void_p), t.ctypes.data_as(C.c_void_p)))

    def step(self):
        self._ck(self.lib.emu_sim_run(self.h))

    def reset(self):
        self._ck(self.lib.emu_sim_reset(self.h))

    def stats(self):
        from mezzo_b200._lib import SimStats
        st = SimStats()
        self.lib.emu_sim_last_stats(self.h, C.byref(st))
        return st.asdict()

    def arrivals(self):
        nt = self.flat.struct.n_trips
        arr = np.empty(max(1, nt))
        self.lib.emu_sim_get_arrivals(self.h, arr.ctypes.data_as(C.c_void_p))
        return arr[:nt]

    def linktimes(self):
        L, P = self.flat.struct.n_links, self.flat.struct.n_periods
        lt = np.empty((L, P))
        self.lib.emu_sim_get_linktimes(self.h, lt.ctypes.data_as(C.c_void_p))
        return lt

    def exit_log(self):
        n = C.c_int64(0)
        self.lib.emu_sim_get_exit_log(self.h, None, 0, C.byref(n))
        ex = np.empty(max(1, n.value), EXIT_DT)
        self.lib.emu_sim_get_exit_log(self.h, ex.ctypes.data_as(C.c_void_p), n.value, C.byref(n))
        return ex[:n.value]

    def close(self):
        if self.h:
            self.lib.emu_sim_destroy(self.h)
            self.h = C.c_void_p()


class EmuPartBackend:
    """Per-rank engine for mezzo_b200.dist.PartitionedSimulation on the CPU: the host emulation, one PROCESS per rank,
    arenas in POSIX shared memory (the stand-in for NVLink peer mappings). Test infrastructure only."""

    def __init__(self, shm_prefix):
        self.prefix = shm_prefix
        self.lib = emu_lib()
        self.lib.emu_sim_create_part.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_char_p, C.POINTER(C.c_void_p)]
        self.lib.emu_sim_attach_peer.argtypes = [C.c_void_p, C.c_int]
        self.lib.emu_sim_min_key.argtypes = [C.c_void_p, C.POINTER(C.c_double), C.POINTER(C.c_double),
                                             C.POINTER(C.c_int32), C.POINTER(C.c_int32)]
        self.lib.emu_sim_final_event.argtypes = [C.c_void_p, C.c_int32, C.c_double]
        self.h = C.c_void_p()

    def _ck(self, rc):
        if rc != 0:
            raise RuntimeError(f"emu error {rc}: {self.lib.emu_last_error().decode()}")

    def create(self, flat, world, rank, node_rank):
        self.flat = flat
        self._nr = np.ascontiguousarray(node_rank, np.int32)
        self._ck(self.lib.emu_sim_create_part(C.byref(flat.struct), world, rank, self._nr.ctypes.data_as(C.c_void_p),
                                              self.prefix.encode() if world > 1 else None, C.byref(self.h)))

    def export_handle(self):
        return self.prefix  # peers map "<prefix><rank>" by name

    def attach(self, peer_rank, handle):
        self._ck(self.lib.emu_sim_attach_peer(self.h, peer_rank))

    def reset(self):
        self._ck(self.lib.emu_sim_reset(self.h))

    def run(self):
        self._ck(self.lib.emu_sim_run(self.h))

    def min_key(self):
        t, tp, sub, node = C.c_double(), C.c_double(), C.c_int32(), C.c_int32()
        self._ck(self.lib.emu_sim_min_key(self.h, C.byref(t), C.byref(tp), C.byref(sub), C.byref(node)))
        return (t.value, tp.value, sub.value, node.value)

    def final_event(self, node, end_time):
        self._ck(self.lib.emu_sim_final_event(self.h, node, end_time))

    def stats(self):
        from mezzo_b200._lib import SimStats
        st = SimStats()
        self.lib.emu_sim_last_stats(self.h, C.byref(st))
        return st.asdict()

    def arrivals(self):
        nt = self.flat.struct.n_trips
        arr = np.empty(max(1, nt))
        self.lib.emu_sim_get_arrivals(self.h, arr.ctypes.data_as(C.c_void_p))
        return arr[:nt]

    def linktimes(self):
        L, P = self.flat.struct.n_links, self.flat.struct.n_periods
        lt = np.empty((L, P))
        self.lib.emu_sim_get_linktimes(self.h, lt.ctypes.data_as(C.c_void_p))
        return lt

    def exit_log(self):
        n = C.c_int64(0)
        self.lib.emu_sim_get_exit_log(self.h, None, 0, C.byref(n))
        ex = np.empty(max(1, n.value), EXIT_DT)
        self.lib.emu_sim_get_exit_log(self.h, ex.ctypes.data_as(C.c_void_p), n.value, C.byref(n))
        return ex[:n.value]

    def destroy(self):
        if self.h:
            self.lib.emu_sim_destroy(self.h)
            self.h = C.c_void_p()
