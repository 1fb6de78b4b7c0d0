"""GPU parity test of the PARTITIONED link update: 2 (or more) GPUs, one process each, arenas mapped through CUDA IPC,
against the sequential CPU oracle. Needs >= 2 visible GPUs (gpurun --gpus 2); skipped otherwise."""
import os
import socket
import sys

import numpy as np
import pytest
import torch.multiprocessing as mp

import mezzo_b200

pytestmark = pytest.mark.gpu


def _worker(rank, world, port, tmp, name, kernel=0):
    sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
    import torch
    import torch.distributed as dist
    import orc
    import simcases
    from mezzo_b200 import dist as mzdist, flatten
    from mezzo_b200.sim import CudaPartBackend
    torch.cuda.set_device(rank)
    dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world)
    try:
        f = flatten.FlatNet(simcases.make(name))
        ps = mzdist.PartitionedSimulation(f, rank, world, CudaPartBackend(device=rank))
        ps.be.set_option(mezzo_b200.MZ_SIM_OPT_KERNEL, kernel)  # 0 by size, 1 a warp per node, 2 a thread per node
        for rep in range(2):  # second pass: Network::reset + rerun must reproduce the first
            if rep:
                ps.reset()
            end = ps.step()
            res = dict(end=end, stats=ps.stats(), arrivals=ps.arrivals(), linktimes=ps.linktimes(), exits=ps.exit_log())
            if rank == 0:
                o = orc.oracle_sim(f)
                oe = np.sort(o["exits"], order=["vid", "entry"])
                assert len(oe) == len(res["exits"])
                assert np.array_equal(oe["vid"], res["exits"]["vid"]) and np.array_equal(oe["link"], res["exits"]["link"])
                if name in simcases.DET:
                    assert np.array_equal(oe["entry"], res["exits"]["entry"]) and np.array_equal(oe["exit"], res["exits"]["exit"])
                    assert np.array_equal(o["arrivals"], res["arrivals"])
                    assert np.array_equal(o["linktimes"], res["linktimes"])
                else:  # N(mu,sd) servers: CUDA log/sin vs glibc, 1e-6 relative (north star)
                    np.testing.assert_allclose(res["exits"]["exit"], oe["exit"], rtol=1e-6, atol=0)
                    np.testing.assert_allclose(res["arrivals"], o["arrivals"], rtol=1e-6, atol=0)
                for k in ("n_traversals", "n_exits", "n_arrived") + (("n_events",) if name in simcases.DET and name not in simcases.SIGNALS else ()):
                    assert o["stats"][k] == res["stats"][k], k
                assert abs(o["stats"]["end_time"] - res["stats"]["end_time"]) <= 1e-6 * o["stats"]["end_time"]
                # the partitioned run must equal the single-GPU run to the last bit (same device arithmetic)
                one = mezzo_b200.Simulation(f, device=0)
                one.step()
                e1 = one.exit_log()
                assert all(np.array_equal(e1[k], res["exits"][k]) for k in ("vid", "link", "entry", "exit"))
                assert np.array_equal(one.arrivals(), res["arrivals"]) and np.array_equal(one.linktimes(), res["linktimes"])
                s1 = one.stats()
                for k in ("n_traversals", "n_exits", "n_arrived", "n_events", "end_time"):
                    assert s1[k] == res["stats"][k], k
                one.close()
        ps.close()
        if rank == 0:
            open(os.path.join(tmp, "ok"), "w").write("1")
    except Exception:
        import traceback
        open(os.path.join(tmp, f"err{rank}"), "w").write(traceback.format_exc())
        raise
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("name,kernel", [("det_ties_small", 0), ("det_two_lane_congested", 0), ("congested_short", 0),
                                         ("two_lane_lookback3", 0), ("det_ties_small", 2), ("rate_changes_det", 2), ("signals_det", 0)])
def test_partitioned_link_update_two_gpus(built, tmp_path, name, kernel):
    if mezzo_b200.lib().mz_device_count() < 2:
        pytest.skip("needs 2 GPUs")
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    try:
        mp.spawn(_worker, args=(2, port, str(tmp_path), name, kernel), nprocs=2, join=True)
    except Exception:
        for r in range(2):
            if (tmp_path / f"err{r}").exists():
                print(f"---- rank {r}\n" + open(tmp_path / f"err{r}").read())
        raise
    assert open(tmp_path / "ok").read() == "1"


@pytest.mark.parametrize("name", ["det_ties_small", "det_two_lane_congested", "signals_det", "incidents_det"])
def test_mezzo_lib_dropin_step_two_gpus_one_process(built, tmp_path, monkeypatch, name):
    """The C++ seam on several GPUs: the UNMODIFIED mezzo_lib executable flow with mezzo_b200/host/NetworkStep.h as its event
    loop and MEZZO_B200_GPUS=2 — one process, an engine per GPU (mz_partition_nodes, mz_sim_create_part,
    mz_sim_attach_peer_local), one host thread per engine. Every file Network::writeall derives from the run, the OD grids and
    the per-vehicle link exits must equal the all-CPU reference run."""
    import filecmp
    sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
    import refrun
    import simcases
    import test_files
    if mezzo_b200.lib().mz_device_count() < 2:
        pytest.skip("needs 2 GPUs")
    if not (refrun.have_ref() and refrun.have_dropin()):
        pytest.skip("oracle/_ref harnesses not built")
    spec = simcases.make(name)
    ref, mine = os.path.join(str(tmp_path), "ref"), os.path.join(str(tmp_path), "dropin2")
    a = refrun.run_reference(spec, workdir=ref, keep=True, save=True)
    monkeypatch.setenv("MEZZO_B200_GPUS", "2")
    b = refrun.run_reference(spec, workdir=mine, keep=True, save=True, binary=refrun.DROPIN_RUN)
    for fn in test_files.OUT_FILES:
        assert filecmp.cmp(os.path.join(ref, fn), os.path.join(mine, fn), shallow=False), fn
    assert np.array_equal(a["arrivals"], b["arrivals"])
    ea, eb = np.sort(a["exits"], order=["vid", "entry"]), np.sort(b["exits"], order=["vid", "entry"])
    assert len(ea) > 1000 and np.array_equal(ea, eb)
    assert a["info"]["currenttime"] == b["info"]["currenttime"]
