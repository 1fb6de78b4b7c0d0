"""CPU tests (no GPU): the C-ABI library loads, exports every symbol include/mezzo_b200.h declares, and refuses to
compute without a device (no CPU fallback)."""
import ctypes as C
import os
import re

import numpy as np
import pytest

import mezzo_b200
from mezzo_b200 import _lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    src = open(os.path.join(ROOT, "include", "mezzo_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(mz_[a-z0-9_]+)\s*\(", src)))


def test_library_exports_every_declared_symbol():
    lib = mezzo_b200.lib()
    syms = declared_symbols()
    assert "mz_sssp_batch" in syms and "mz_graph_create" in syms
    for s in syms:
        assert hasattr(lib, s), f"{s} declared in include/mezzo_b200.h but not exported"
    assert lib.mz_version() >= 100


def test_no_cpu_fallback_without_device():
    lib = mezzo_b200.lib()
    if lib.mz_device_count() > 0:
        pytest.skip("a CUDA device is present")
    with pytest.raises(mezzo_b200.MzError) as e:
        mezzo_b200.Graph.from_arrays(3, 2, [0, 1], [1, 2])
    assert e.value.code == _lib.MZ_ENODEV


def test_argument_validation_without_device():
    lib = mezzo_b200.lib()
    h = C.c_void_p()
    up = np.array([0, 1], np.int32)
    assert lib.mz_graph_create(0, 0, 2, 0, None, up.ctypes.data_as(C.c_void_p), up.ctypes.data_as(C.c_void_p), None,
                               C.byref(h)) == _lib.MZ_EINVAL
    assert lib.mz_graph_create(0, 3, 2, 0, None, None, None, None, C.byref(h)) == _lib.MZ_EINVAL
    assert b"" != lib.mz_last_error()
    assert lib.mz_sssp_batch(None, 1, None, None, 0, None, None, None) == _lib.MZ_EINVAL
    assert lib.mz_graph_destroy(None) == _lib.MZ_OK


def test_vtypes_draw_argument_validation_and_known_answer():
    """mz_vtypes_draw (host only): bad arguments are refused; the draw is Park-Miller (A = 48271, M = 2^31 - 1) from the
    given seed, first class whose cumulative probability reaches the draw (B/vtypes.h:61-71, B/Random.cpp:85-98)."""
    lib = mezzo_b200.lib()
    lib.mz_vtypes_draw.argtypes = [C.c_int64, C.c_int32, C.c_void_p, C.c_int64, C.c_void_p]
    prob = np.array([0.5, 0.3, 0.2])
    out = np.full(8, -1, np.int32)
    ptr = lambda a: a.ctypes.data_as(C.c_void_p)  # noqa: E731
    assert lib.mz_vtypes_draw(42, 0, ptr(prob), 8, ptr(out)) == _lib.MZ_EINVAL
    assert lib.mz_vtypes_draw(0, 3, ptr(prob), 8, ptr(out)) == _lib.MZ_EINVAL  # seed 0 = clock seeding in the reference
    assert lib.mz_vtypes_draw(42, 3, None, 8, ptr(out)) == _lib.MZ_EINVAL
    assert lib.mz_vtypes_draw(42, 3, ptr(prob), 8, ptr(out)) == _lib.MZ_OK
    seed, want = 42, []
    for _ in range(8):
        seed = 48271 * (seed % 44488) - 3399 * (seed // 44488)
        if seed <= 0:
            seed += 2147483647
        r = seed / 2147483647
        want.append(0 if 0.5 >= r else 1 if 0.5 + 0.3 >= r else 2)
    assert out.tolist() == want
    # rounding may leave the cumulative sum below the draw: the last class then (the reference would return NULL)
    assert lib.mz_vtypes_draw(42, 2, ptr(np.array([1e-9, 1e-9])), 8, ptr(out)) == _lib.MZ_OK and set(out.tolist()) == {1}


def test_host_mirror_prohibitor_semantics():
    """set_turning_prohibitor follows the char XOR of B/Graph.cpp:709-717 (checked against the golden legal masks
    in test_oracle_sssp); a repeated prohibitor is an error like the reference's assert."""
    g = mezzo_b200.Graph(4, 3)
    g.addLink(0, 0, 1)
    g.addLink(1, 1, 2)
    g.addLink(2, 1, 3)
    g.set_downlink_indices()
    g.set_turning_prohibitor(0, 2)
    assert g.legal[0] == 0xFD
    with pytest.raises(mezzo_b200.MzError):
        g.set_turning_prohibitor(0, 2)


def test_shared_stochastic_server_is_named_not_simulated():
    """a stochastic server shared by several nodes has ONE random stream consumed in global event order; the
    host-side check names the server ids and their nodes, and the engine refuses (no approximation)."""
    import mezzo_b200
    from mezzo_b200 import flatten, synth
    spec = synth.sim_grid(nx=4, ny=4, seed=1, od_every=2, n_od=4, trips_per_hour=100, horizon=600.0, shared_server=True)
    f = flatten.FlatNet(spec)
    assert list(f.shared_stochastic_servers) == [0] and len(f.shared_stochastic_servers[0]) == 16
    with pytest.raises(mezzo_b200._lib.MzError) as e:
        mezzo_b200.Simulation(f)
    assert e.value.code == mezzo_b200._lib.MZ_ELIMIT and "shared by several nodes: 0 (nodes" in str(e.value)
    det = synth.sim_grid(nx=4, ny=4, seed=1, od_every=2, n_od=4, trips_per_hour=100, horizon=600.0, shared_server=True,
                         server_type=2, server_mu=1.5, server_sd=0.0)
    assert flatten.FlatNet(det).shared_stochastic_servers == {}  # deterministic servers may be shared freely
