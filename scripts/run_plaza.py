#!/usr/bin/env python
"""matlab/PlazaPose2.m without MATLAB: load a Plaza .mat, build the script's graph, optimise with LM (or GN) on the B200 engine
(default) or the CPU oracle (--oracle), print the script's summary numbers.

  python scripts/run_plaza.py /path/to/Plaza2.mat [--linear] [--gn] [--oracle]
"""
import argparse
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from gpslam_b200 import datasets as ds  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("mat")
ap.add_argument("--linear", action="store_true", help="useLinearPose2 (matlab/PlazaPose2.m:27)")
ap.add_argument("--gn", action="store_true", help="useGaussNewton (:35)")
ap.add_argument("--oracle", action="store_true", help="solve on the CPU oracle instead of the engine (test infrastructure)")
a = ap.parse_args()
d = ds.load_plaza(a.mat)
if a.oracle:
    from oracle import pyoracle as po
    make = lambda grp, n, l: po.Graph(grp, n, l)
else:
    import gpslam_b200 as gb
    make = lambda grp, n, l: gb.Graph(grp, n, l)
g, info = ds.build_plaza(d, make, use_linear=a.linear)
e0 = g.error()
t0 = time.perf_counter()
st = g.optimize(use_lm=not a.gn)
dt = time.perf_counter() - t0
P, _, _ = g.get_values()
pos, rot = ds.plaza_errors(d, P)
print(json.dumps({"poses": info["n_poses"], "ranges_used": info["n_ranges_used"], "outliers": info["n_outliers"], "init_error": e0, "final_error": st.error_final,
                  "iterations": st.iterations, "avg_iteration_s": dt / max(1, st.iterations), "avg_position_error_m": pos, "avg_rotation_error_rad": rot,
                  "solver": "oracle (CPU)" if a.oracle else "engine (B200)"}))
