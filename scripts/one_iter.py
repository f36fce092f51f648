#!/usr/bin/env python
"""A few Gauss-Newton iterations of one config, for profiler runs (ncu launch lists / --set full captures of single kernels).
  python scripts/one_iter.py [config] [states] [iterations]"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import gpslam_b200 as gb  # noqa: E402
from gpslam_b200 import synth  # noqa: E402

name = sys.argv[1] if len(sys.argv) > 1 else "C3"
cfg = synth.config(name)
if len(sys.argv) > 2 and int(sys.argv[2]):
    cfg.n_states = int(sys.argv[2])
    if cfg.n_closures:
        cfg.closure_min_gap = min(cfg.closure_min_gap, cfg.n_states // 10)
g, _ = synth.build(cfg, lambda grp, n, l: gb.Graph(grp, n, l))
g.linearize()
st = g.optimize(n_iter=int(sys.argv[3]) if len(sys.argv) > 3 else 3, use_lm=False)
print("%s %d states: %.3f ms / iteration, error %.6e" % (name, cfg.n_states, st.total_ms / st.iterations, st.error_final))
