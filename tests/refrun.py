"""TEST INFRASTRUCTURE: drive the compiled UNMODIFIED reference (oracle/_ref/mezzo_ref_run) on a NetSpec.

Writes the reference's own input files (grammars: SURVEY.md §5f) into a scratch directory, runs the harness
(B/main.cpp call sequence, timing only Network::step) and loads the binary dumps."""
import os
import subprocess
import tempfile

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF_RUN = os.path.join(ROOT, "oracle", "_ref", "mezzo_ref_run")
# the same harness + reference library with Network::step's event loop replaced by the GPU engine (NetworkStep.h)
DROPIN_RUN = os.path.join(ROOT, "oracle", "_ref", "mezzo_dropin_run")

EXIT_DT = np.dtype([("vid", "<i4"), ("lid", "<i4"), ("entry", "<f8"), ("exit", "<f8")])
TRIP_DT = np.dtype([("vid", "<i4"), ("route", "<i4"), ("oid", "<i4"), ("did", "<i4"), ("time", "<f8"), ("length", "<f8")])

MASTER = """#input_files
network= net.dat
turnings= turnings.dat
signals= signal.dat
histtimes= histtimes.dat
routes= routes.dat
demand= demand.dat
incident= noincident.dat
vehicletypes= vehiclemix.dat
virtuallinks= virtuallinks.dat
serverrates= serverrates.dat
#output_files
linktimes= output/linktimes.dat
output= output/output.dat
summary= output/summary.dat
speeds= output/speeds.dat
inflows= output/inflows.dat
outflows= output/outflows.dat
queuelengths= output/queuelengths.dat
densities= output/densities.dat
#scenario
starttime= {starttime:g}
stoptime= {stoptime:g}
calc_paths= {calc_paths}
traveltime_alpha= 0.4
parameters= parameters.dat
nobackground= ----
"""

# key order is fixed by Parameters::read_parameters (B/parameters.cpp:118+)
PARAMS = """#drawing_parameters
   draw_link_ids= 0
   link_thickness= 1
   node_thickness= 1
   node_radius= 3
   queue_thickness= 6
   selected_thickness= 10
   show_background_image= 1
   linkcolor= black
   nodecolor= black
   queuecolor= red
   backgroundcolor= white
   selectedcolor= green
   gui_update_step= 0.2
#moe_parameters
   moe_speed_update= 900.0
   moe_inflow_update= 900.0
   moe_outflow_update= 900.0
   moe_queue_update= 900.0
   moe_density_update= 900.0
   linktime_alpha= 0.2
#assignment_matrix_parameters
   use_ass_matrix= 0
   ass_link_period= 900.0
   ass_od_period= 900.0
#turning_parameters
   default_lookback_size= {default_lookback_size}
   turn_penalty_cost= 99999.0
#server_parameters
   od_servers_deterministic= {od_servers_deterministic}
   odserver_sigma= 0.2
   sd_server_scale= 1.0
   server_type= {server_type}
#vehicle_parameters
   standard_veh_length= {standard_veh_length}
#route_parameters
   update_interval_routes= {update_interval_routes}
   mnl_theta= -0.00417
   kirchoff_alpha= -1.0
   delete_bad_routes= {delete_bad_routes}
   max_rel_route_cost= {max_rel_route_cost}
   small_od_rate= {small_od_rate}
#mime_parameters
   mime_comm_step= 0.4
   mime_min_queue_length= 20
   mime_queue_dis_speed= 6
   vissim_step= 0.1
   sim_speed_factor= 1.9
#transit_demand_parameters
   demand_format= 1
   demand_scale= 1.0
#transit_control_parameters
   riding_time_weight= 1.0
   dwell_time_weight= 1.0
   waiting_time_weight= 1.0
   holding_time_weight= 1.0
   compliance_rate= 0.8
   transfer_sync= 0
#day2day_assignment
    default_alpha_RTI= 0.7
"""


def _fmt(v):
    return repr(float(v)) if isinstance(v, float) else str(v)


def write_inputs(spec, d, calc_paths=0):
    os.makedirs(os.path.join(d, "output"), exist_ok=True)
    w = lambda name, text: open(os.path.join(d, name), "w").write(text)  # noqa: E731
    w("masterfile.mezzo", MASTER.format(starttime=spec["starttime"], stoptime=spec["stoptime"], calc_paths=calc_paths))
    w("parameters.dat", PARAMS.format(**dict(dict(delete_bad_routes=0, max_rel_route_cost=1.5, small_od_rate=0.0),
                                             **spec["params"])))
    out = [f"servers: {len(spec['servers'])}"]
    out += ["{ %d %d %s %s %s }" % (s[0], s[1], _fmt(float(s[2])), _fmt(float(s[3])), _fmt(float(s[4])))
            for s in spec["servers"]]
    out.append(f"nodes: {len(spec['nodes'])}")
    for n in spec["nodes"]:
        out.append("{ %d %d %s %s %s}" % (n[0], n[1], _fmt(float(n[2])), _fmt(float(n[3])),
                                          (str(n[4]) + " ") if n[1] == 2 else ""))
    out.append(f"sdfuncs: {len(spec['sdfuncs'])}")
    out += ["{ " + " ".join(_fmt(v) for v in s) + " }" for s in spec["sdfuncs"]]
    out.append(f"links: {len(spec['links'])}")
    out += ["{ %d %d %d %d %d %d }" % tuple(l) for l in spec["links"]]
    w("net.dat", "\n".join(out) + "\n")
    w("turnings.dat", f"turnings: {len(spec['turnings'])}\n" +
      "\n".join("{ %d %d %d %d %d %d }" % tuple(t) for t in spec["turnings"]) + "\n\ngiveways: 0\n")
    w("demand.dat", f"od_pairs: {len(spec['ods'])}\nscale: 1.0\n" +
      "\n".join("{ %d %d %s }" % (o, dd, _fmt(float(r))) for o, dd, r in spec["ods"]) +
      f"\nslices: {len(spec.get('slices') or [])}\n" +
      "".join(f"od_pairs: {len(rates)}\nscale: {_fmt(float(scale))}\nloadtime: {_fmt(float(lt))}\n" +
              "".join("{ %d %d %d }\n" % (o, dd, int(r)) for o, dd, r in rates)
              for lt, scale, rates in (spec.get("slices") or [])))
    w("routes.dat", f"routes: {len(spec['routes'])}\n" +
      "\n".join("{ %d %d %d %d{ %s} }" % (r[0], r[1], r[2], len(r[3]), " ".join(map(str, r[3]))) for r in spec["routes"])
      + "\n")
    if spec.get("histtimes") is None:
        w("histtimes.dat", f"links: 0\nperiods: {spec['n_periods']}\nperiodlength: {_fmt(spec['periodlength'])}\n")
    else:
        h = spec["histtimes"]  # dict lid -> list of P floats
        w("histtimes.dat", f"links: {len(h)}\nperiods: {spec['n_periods']}\nperiodlength: {_fmt(spec['periodlength'])}\n" +
          "\n".join("{ %d %s }" % (lid, " ".join(repr(float(v)) for v in row)) for lid, row in h.items()) + "\n")
    w("vehiclemix.dat", f"vtypes: {len(spec['vtypes'])}\n" +
      "\n".join("{ %d %s %s %s }" % (v[0], v[1], _fmt(float(v[2])), _fmt(float(v[3]))) for v in spec["vtypes"]) + "\n")
    rates = spec.get("serverrates") or []
    w("serverrates.dat", f"rates: {len(rates)}\n" +
      "".join("{ %d %s %s %s }\n" % (r[0], _fmt(float(r[1])), _fmt(float(r[2])), _fmt(float(r[3]))) for r in rates))
    sig = spec.get("signals") or []
    txt = [f"controls: {len(sig)}"]
    for cid, plans in sig:
        txt.append("{ %d %d" % (cid, len(plans)))
        for pid, start, stop, offset, cycle, stages in plans:
            txt.append("  { %d %s %s %s %s %d" % (pid, _fmt(float(start)), _fmt(float(stop)), _fmt(float(offset)),
                                                 _fmt(float(cycle)), len(stages)))
            for sid, sstart, dur, turns in stages:
                txt.append("    { %d %s %s %d { %s } }" % (sid, _fmt(float(sstart)), _fmt(float(dur)), len(turns),
                                                          " ".join(map(str, turns))))
            txt.append("  }")
        txt.append("}")
    w("signal.dat", "\n".join(txt) + "\n")
    w("virtuallinks.dat", "virtuallinks: 0\n")
    isd, inc = spec.get("incident_sdfuncs") or [], spec.get("incidents") or []
    w("noincident.dat", f"sdfuncs: {len(isd)}\n" + "".join("{ " + " ".join(_fmt(v) for v in s) + " }\n" for s in isd) +
      f"incidents: {len(inc)}\n" +
      "".join("{ %d %d %s %s %s %s %s %d }\n" % (x[0], x[1], _fmt(float(x[2])), _fmt(float(x[3])), _fmt(float(x[4])),
                                                _fmt(float(x[5])), _fmt(float(x[6])), int(x[7])) for x in inc) +
      "parameters: 4\n{1.0 0.1}\n{1.0 0.1}\n{1.0 0.1}\n{1.0 0.1}\nX1:\n{2.0 0.5}\n")
    w("assign_links.dat", "no_obs_links: 0\n{ }\n")
    for t in ("transit_routes.dat", "transit_network.dat", "transit_fleet.dat", "transit_demand.dat"):
        w(t, "none: 0\n")


def have_ref():
    return os.path.exists(REF_RUN)


def have_dropin():
    return os.path.exists(DROPIN_RUN)


def run_reference(spec, workdir=None, calc_paths=0, keep=False, timeout=3600, save=False, binary=None):
    """Returns dict(exits, trips, arrivals [n,9], info{}, workdir). save=True also runs NetworkThread::saveresults
    (Network::writeall): output/linktimes.dat[.clean], output.dat, summary.dat, MOE files in the work directory."""
    assert have_ref(), "oracle/_ref/mezzo_ref_run missing (make -C oracle ref, needs /root/reference)"
    tmp = None
    if workdir is None:
        tmp = tempfile.TemporaryDirectory(prefix="mzref_")
        workdir = tmp.name
    write_inputs(spec, workdir, calc_paths)
    os.makedirs(os.path.join(workdir, "mz"), exist_ok=True)
    with open(os.path.join(workdir, "mz", "stdout.log"), "w") as so:
        subprocess.run([binary or REF_RUN, "masterfile.mezzo", str(spec["randseed"]), "mz/"] + (["save"] if save else []),
                       cwd=workdir, stdout=so,
                       stderr=subprocess.STDOUT, check=True, timeout=timeout)
    res = dict(
        exits=np.fromfile(os.path.join(workdir, "mz", "exits.bin"), EXIT_DT),
        trips=np.fromfile(os.path.join(workdir, "mz", "trips.bin"), TRIP_DT),
        arrivals=np.fromfile(os.path.join(workdir, "mz", "arrivals.bin"), "<f8").reshape(-1, 9),
        info=dict(l.strip().split("=") for l in open(os.path.join(workdir, "mz", "info.txt"))),
        workdir=workdir)
    if tmp is not None and not keep:
        tmp.cleanup()
    return res
