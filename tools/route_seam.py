"""Measurement (GPU box): route generation through the REAL mezzo_lib seam. The unmodified reference executable flow
(oracle/_ref/mezzo_ref_run, calc_paths=1: Network::init_shortest_path + shortest_paths_all on the CPU) against the same flow
with mezzo_b200/host/Graph.h force-included (oracle/_ref/mezzo_dropin_run: every labelCorrecting call of the unpatched
Network::shortest_paths_all is answered by the GPU engine, batched by the drop-in's prefetch). Compares the routes.dat both
write back and reports the time of NetworkThread::init (readers + route generation) of each.
usage: python tools/route_seam.py [out.json]"""
import json, os, sys, tempfile, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import refrun
from mezzo_b200 import synth
import test_fullsize_gpu as tf

spec = synth.sim_grid(**dict(tf.CONFIG3, total_trips=20000))
spec["stoptime"] = 2700.0  # two re-runs of the route search (runtime / update_interval_routes - 1)
out = {"workload": "config 3 network (%d links, %d OD pairs, no routes given), calc_paths=1" % (len(spec["links"]), len(spec["ods"]))}
res = {}
for name, binary, env in (("reference_cpu", refrun.REF_RUN, {}), ("dropin_gpu", refrun.DROPIN_RUN, {}),
                          ("dropin_gpu_no_prefetch", refrun.DROPIN_RUN, {"MEZZO_B200_PREFETCH": "0"})):
    d = tempfile.mkdtemp(prefix="mzseam_")
    os.environ.pop("MEZZO_B200_PREFETCH", None)
    os.environ.update(env)
    t0 = time.time()
    r = refrun.run_reference(spec, workdir=d, calc_paths=1, keep=True, binary=binary, timeout=3000)
    res[name] = dict(init_seconds=float(r["info"]["init_seconds"]), wall_seconds=time.time() - t0,
                     routes=open(os.path.join(d, "routes.dat")).read(),
                     prefetch_batches=int(r["info"].get("sssp_prefetch_batches", 0)),
                     prefetch_served=int(r["info"].get("sssp_prefetch_served", 0)))
    print(name, {k: v for k, v in res[name].items() if k != "routes"}, flush=True)
same = res["reference_cpu"]["routes"] == res["dropin_gpu"]["routes"] == res["dropin_gpu_no_prefetch"]["routes"]
n_routes = res["reference_cpu"]["routes"].count("\n") - 1
for k in res:
    out[k] = {a: b for a, b in res[k].items() if a != "routes"}
out["routes_identical"] = same
out["n_routes"] = n_routes
out["speedup_init"] = res["reference_cpu"]["init_seconds"] / res["dropin_gpu"]["init_seconds"]
print(json.dumps(out))
if len(sys.argv) > 1:
    json.dump(out, open(sys.argv[1], "w"), indent=1)
assert same, "routes.dat differs between the CPU reference and the GPU drop-in"
