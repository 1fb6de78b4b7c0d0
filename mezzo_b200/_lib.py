"""ctypes loader for the in-tree C-ABI library (mezzo_b200/libmezzo_b200.so, built by mezzo_b200/csrc/Makefile).

There is no Python or CPU implementation behind these bindings: if the library is missing the import fails loudly."""
import ctypes as C
import os

HERE = os.path.dirname(os.path.abspath(__file__))
# MZ_LIB: tuning aid — load another build of the same library (e.g. a different MZ_PTHREADS), never a fallback
LIB_PATH = os.environ.get("MZ_LIB") or os.path.join(HERE, "libmezzo_b200.so")

MZ_OK, MZ_EINVAL, MZ_ENODEV, MZ_ECUDA, MZ_ENOMEM, MZ_ESTATE, MZ_ELIMIT = 0, -1, -2, -3, -4, -5, -6
MZ_SSSP_EXACT, MZ_SSSP_AUTO = 0, 1
MZ_SIM_OPT_MAX_EVENTS, MZ_SIM_OPT_CHECK_EVERY, MZ_SIM_OPT_BLOCKS, MZ_SIM_OPT_KERNEL = 1, 2, 3, 4
MZ_OPT_CHUNK_TREES, MZ_OPT_COUNT_CANONICAL, MZ_OPT_DELTA_MILLI, MZ_OPT_MAX_STEPS = 1, 2, 3, 4
SP_UNSET, SP_START = -2, -1
SP_INFINITY = 9999999.0


class MzError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__(f"mezzo_b200 error {code}: {msg}")
        self.code = code


class SsspStats(C.Structure):
    _fields_ = [("canonical_relax", C.c_int64), ("evaluated_relax", C.c_int64),
                ("reached_links", C.c_int64), ("kernel_ms", C.c_double),
                ("main_kernel_ms", C.c_double), ("total_ms", C.c_double), ("n_exact_redo", C.c_int32),
                ("n_launches", C.c_int32), ("n_chunks", C.c_int32), ("main_kernel_launches", C.c_int32),
                ("fast_steps", C.c_int64)]

    def asdict(self):
        return {k: getattr(self, k) for k, _ in self._fields_}


class MzNetStruct(C.Structure):
    """MzNet of include/mezzo_b200.h (field order is the ABI)."""
    _i, _p = C.c_int32, C.c_void_p
    _fields_ = ([(n, C.c_int32) for n in ("n_nodes", "n_links", "n_sdfuncs", "n_servers", "n_turnings", "n_dests",
                                          "n_origins", "n_routes", "n_trips", "n_periods", "n_odpairs",
                                          "standard_veh_length", "server_type", "use_giveway")] +
                [("randseed", C.c_int64), ("runtime", C.c_double), ("periodlength", C.c_double)] +
                [(n, C.c_void_p) for n in ("link_length", "link_lanes", "link_sdfunc", "link_in_node", "link_out_node",
                                           "link_hist", "sd_type", "sd_vfree", "sd_vmin", "sd_romax", "sd_romin",
                                           "sd_alpha", "sd_beta", "srv_type", "srv_mu", "srv_sd", "srv_delay",
                                           "turn_node", "turn_server", "turn_in", "turn_out", "turn_lookback",
                                           "dest_node", "dest_server", "origin_node", "route_off", "route_links",
                                           "trip_time", "trip_origin", "trip_route", "trip_length", "trip_od",
                                           "trip_prev_time")] +
                [("n_rate_changes", C.c_int32)] +
                [(n, C.c_void_p) for n in ("chg_server", "chg_time", "chg_mu", "chg_sd")] +
                [(n, C.c_int32) for n in ("n_sig_controls", "n_sig_plans", "n_sig_stages")] +
                [(n, C.c_void_p) for n in ("sigc_plan_off", "sigp_start", "sigp_stop", "sigp_stage_off", "sigs_duration",
                                           "sigs_turn_off", "sigs_turns")] +
                [("phantom_final_t", C.c_double), ("phantom_final_tp", C.c_double), ("n_incidents", C.c_int32)] +
                [(n, C.c_void_p) for n in ("inc_link", "inc_sdfunc", "inc_start", "inc_stop")] +
                [(n, C.c_void_p) for n in ("trip_stop_off", "stop_link", "stop_pos")] +
                [("n_transit_init_steps", C.c_int32)])


class SimStats(C.Structure):
    _fields_ = [("n_traversals", C.c_int64), ("n_exits", C.c_int64), ("n_events", C.c_int64),
                ("n_arrived", C.c_int64), ("n_supersteps", C.c_int64), ("n_launches", C.c_int64),
                ("kernel_ms", C.c_double), ("total_ms", C.c_double), ("end_time", C.c_double),
                ("n_stop_visits", C.c_int64), ("n_bus_refused", C.c_int64)]

    def asdict(self):
        return {k: getattr(self, k) for k, _ in self._fields_}


class StopVisit(C.Structure):
    """mz_stop_visit of include/mezzo_b200.h."""
    _fields_ = [("veh", C.c_int32), ("stop", C.c_int32), ("link", C.c_int32), ("pad", C.c_int32),
                ("t_enter", C.c_double), ("t_arrive", C.c_double)]


_lib = None


def lib():
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ImportError(f"{LIB_PATH} not built: run `make -C mezzo_b200/csrc` (or __graft_entry__.build()); "
                          "mezzo_b200 has no CPU fallback")
    l = C.CDLL(LIB_PATH)
    vp, i32, i64, dbl = C.c_void_p, C.c_int32, C.c_int64, C.c_double
    l.mz_last_error.restype = C.c_char_p
    l.mz_version.restype = C.c_int
    l.mz_device_count.restype = C.c_int
    l.mz_graph_create.argtypes = [C.c_int, i32, i32, i32, vp, vp, vp, vp, C.POINTER(vp)]
    l.mz_linktimes_set.argtypes = [vp, i32, dbl, vp]
    l.mz_sssp_batch.argtypes = [vp, i32, vp, vp, C.c_int, vp, vp, vp]
    l.mz_sssp_routes.argtypes = [vp, i32, vp, vp, C.c_int, vp, vp, vp, vp, i64, C.POINTER(i64)]
    l.mz_sssp_last_stats.argtypes = [vp, C.POINTER(SsspStats)]
    l.mz_graph_set_option.argtypes = [vp, C.c_int, i64]
    l.mz_graph_set_stream.argtypes = [vp, vp]
    l.mz_graph_destroy.argtypes = [vp]
    _lib = l
    return l


def check(rc):
    if rc != MZ_OK:
        raise MzError(rc, lib().mz_last_error().decode())
