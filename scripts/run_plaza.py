#!/usr/bin/env python
"""matlab/PlazaPose2.m without MATLAB: load a Plaza .mat, build the script's graph, optimise with LM (or GN) on the B200 engine,
print the script's summary numbers.  (The same graph on the CPU oracle: tests/test_datasets.py.)

  python scripts/run_plaza.py /path/to/Plaza2.mat [--linear] [--gn]
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
a = ap.parse_args()
d = ds.load_plaza(a.mat)
import gpslam_b200 as gb  # noqa: E402
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
                  "solver": "engine (B200)"}))
