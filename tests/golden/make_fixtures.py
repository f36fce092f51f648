#!/usr/bin/env python
"""Generates the committed golden fixtures of tests/golden/ from the CPU oracle (oracle/, the restated reference).

  python tests/golden/make_fixtures.py        # rewrites tests/golden/graph_*.npz

The real reference cannot run (its every translation unit needs GTSAM/Eigen/Boost, absent here: SURVEY.md §8c), so the fixtures
come from the oracle, which is itself pinned by the reference's own unit-test vectors (tests/test_oracle_golden.py).  They are
small seeded graphs, one per factor family / state manifold; for each: every factor's whitened JacobianFactor [A|b], the graph
error, the values after one Gauss-Newton iteration and after LM convergence.  tests/test_golden_fixtures.py checks the oracle
(CPU) and the CUDA engine (GPU) against them, so neither side can drift unnoticed.
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)

from gpslam_b200 import synth  # noqa: E402

POSE3, POSE2, ROT3, LINEAR = 0, 1, 2, 3

# name -> (synthetic config, size overrides); "linear" uses the hand-built 2DLinear graph of tests/test_gpu_parity.py
CASES = {
    "pose3": ("C3", dict(n_states=12, n_landmarks=3, prior_every=5, range_per_state=0.8)),
    "pose3_gps_proj": ("C3", dict(n_states=14, n_landmarks=4, prior_every=5, range_per_state=0.4, gps_every=3, proj_per_state=0.5)),
    "pose3_loops": ("C5", dict(n_states=14, n_landmarks=2, prior_every=6, range_per_state=0.6, n_closures=2, closure_min_gap=5, closure_ends=True)),
    "pose3vw": ("VW", dict(n_states=13, prior_every=5, gps_every=2, n_closures=1, closure_min_gap=6)),
    "pose2": ("C1", dict(n_states=12)),
    "pose2_loops": ("C1", dict(n_states=13, n_closures=2, closure_min_gap=4)),
    "rot3": ("C4", dict(n_states=13)),
    "linear": (None, dict(n=12)),
}


def build(name, make):
    cfg_name, kw = CASES[name]
    if cfg_name is None:
        from tests.test_gpu_parity import linear_graph
        return linear_graph(make, **kw)
    cfg = synth.config(cfg_name)
    for k, v in kw.items():
        setattr(cfg, k, v)
    g, _ = synth.build(cfg, make)
    return g


def flatten_factor(A, b):
    return np.concatenate([a.ravel(order="F") for a in A] + [np.asarray(b).ravel()])


def generate(name):
    from oracle import pyoracle as po
    o = build(name, lambda grp, n, l: po.Graph(grp, n, l))
    out = {"error0": np.array([o.error()])}
    nf = o.num_factors()
    Ab, off = [], [0]
    for k in range(nf):
        A, b = o.linearize_factor(k)
        Ab.append(flatten_factor(A, b)); off.append(off[-1] + len(Ab[-1]))
    out["Ab"] = np.concatenate(Ab); out["Ab_off"] = np.array(off, dtype=np.int64)
    st = o.optimize(n_iter=1, use_lm=False)
    P, V, L = o.get_values()
    out.update(P1=P, V1=V, L1=L, error1=np.array([st.error_final]))
    st = o.optimize(use_lm=True)
    P, V, L = o.get_values()
    out.update(Pc=P, Vc=V, Lc=L, errorc=np.array([st.error_final]), iters=np.array([st.iterations + 0]))
    return out


if __name__ == "__main__":
    for name in (sys.argv[1:] or CASES):  # optional: only the named cases (existing fixtures stay byte-identical)
        d = generate(name)
        np.savez_compressed(os.path.join(HERE, "graph_%s.npz" % name), **d)
        print(name, "factors", len(d["Ab_off"]) - 1, "error0 %.6e -> %.6e (LM %d iterations after one GN step)" % (d["error0"][0], d["errorc"][0], d["iters"][0]))
