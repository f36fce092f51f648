#!/usr/bin/env python
"""Run one of BASELINE.json's configs (C1..C5, optionally resized) on the CUDA engine and print one JSON line with the
Gauss-Newton iteration rate, the per-stage device times and the solver geometry.  Not the bench (bench.py measures C3);
this records the other configs next to it.

  python scripts/run_config.py --config C5 --states 1000000 --steps 10
  python scripts/run_config.py --config C4 --steps 10 --lm         # LM to convergence afterwards, reports iterations
"""
import argparse
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--config", default="C5")
    ap.add_argument("--states", type=int, default=0)
    ap.add_argument("--closures", type=int, default=-1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--lm", action="store_true", help="after the timed GN iterations, run LM to convergence and report it")
    args = ap.parse_args()
    import numpy as np
    import gpslam_b200 as gb
    from gpslam_b200 import synth
    cfg = synth.config(args.config)
    if args.states:
        cfg.n_states = args.states
    if args.closures >= 0:
        cfg.n_closures = args.closures
    if cfg.n_closures and cfg.closure_min_gap >= cfg.n_states // 2:
        cfg.closure_min_gap = cfg.n_states // 10
    t0 = time.perf_counter()
    g, _ = synth.build(cfg, lambda grp, n, l: gb.Graph(grp, n, l))
    build_s = time.perf_counter() - t0
    sz = g.sizes()
    err0 = g.linearize()
    out = {"config": args.config, "states": cfg.n_states, "closures": cfg.n_closures, "landmark_dims": sz.border_dim, "gp_factors": sz.n_gp,
           "other_factors": sz.n_extra, "solver_levels": sz.levels, "hbm_resident_mb": sz.hbm_bytes / 1e6, "graph_build_s": build_s, "error_initial": err0}
    g.optimize(n_iter=max(args.warmup, 1), use_lm=False)
    st = g.optimize(n_iter=args.steps, use_lm=False)
    out.update({"gn_iterations_per_s": 1e3 * args.steps / st.total_ms, "ms_per_iteration": st.total_ms / args.steps, "launches_per_iteration": g.launches() / args.steps,
                "error_after": st.error_final})
    out["stages_ms"] = {n: g.time_stage(k, 10) for k, n in ((0, "linearise_gp"), (1, "linearise_other"), (2, "assemble"), (3, "solve"), (4, "retract"), (5, "solve_fwd_level0"))}
    if args.lm:
        t0 = time.perf_counter()
        st = g.optimize(use_lm=True)
        out["lm"] = {"iterations": st.iterations, "error_final": st.error_final, "status": st.status, "seconds": time.perf_counter() - t0}
    print(json.dumps(out))


if __name__ == "__main__":
    main()
