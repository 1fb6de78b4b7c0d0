"""Generates tests/golden/sim_golden.npz by running the UNMODIFIED reference (oracle/_ref/mezzo_ref_run = every
B/*.cpp compiled in place, exit log taken by the compile-time macro of oracle/ref_hook_link.h) on the seeded networks
of tests/simcases.py. Authoring container only:

    make -C oracle ref && python tests/golden/make_sim_golden.py
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
import refrun  # noqa: E402
import simcases  # noqa: E402


def main():
    out = {}
    for name in simcases.GOLDEN:
        r = refrun.run_reference(simcases.make(name))
        ex = np.sort(r["exits"], order=["vid", "entry"])
        out[f"{name}_vid"] = ex["vid"]
        out[f"{name}_lid"] = ex["lid"]
        out[f"{name}_entry"] = ex["entry"]
        out[f"{name}_exit"] = ex["exit"]
        out[f"{name}_arrivals"] = r["arrivals"]
        out[f"{name}_trip_time"] = r["trips"]["time"]
        out[f"{name}_trip_route"] = r["trips"]["route"]
        if len(simcases.make(name)["vtypes"]) > 1:  # the class drawn per trip (Vtypes::random_vtype), seen through its length
            out[f"{name}_trip_length"] = r["trips"]["length"]
        out[f"{name}_currenttime"] = np.array(float(r["info"]["currenttime"]))
        print(name, len(ex), "exits", len(r["arrivals"]), "arrivals")
    np.savez_compressed(os.path.join(HERE, "sim_golden.npz"), **out)


if __name__ == "__main__":
    main()
