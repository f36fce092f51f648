#!/usr/bin/env python
"""A/B timing of engine variants selected by environment switches (read at gpb_graph_finalize), on config C3, one GPU.

  python scripts/ab_variants.py "GPB_LIN_VARIANT=0" "GPB_LIN_VARIANT=1" ...       # each argument: space-separated VAR=VALUE list

Per variant: per-stage device times (gpb_time_stage) and ms per Gauss-Newton iteration over 20 iterations.  A tuning aid, not a
bench value."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import gpslam_b200 as gb  # noqa: E402
from gpslam_b200 import synth  # noqa: E402

STAGES = ((0, "lin_gp"), (1, "lin_other"), (2, "assemble"), (3, "solve"), (4, "retract"), (5, "fwd0"), (6, "spine0"), (7, "panel0"), (8, "bwd"))


def main():
    name = os.environ.get("AB_CONFIG", "C3")
    for spec in sys.argv[1:] or [""]:
        sets = dict(kv.split("=", 1) for kv in spec.split() if "=" in kv)
        for k, v in sets.items():
            os.environ[k] = v
        cfg = synth.config(name)
        if os.environ.get("AB_STATES"):
            cfg.n_states = int(os.environ["AB_STATES"])
        g, _ = synth.build(cfg, lambda grp, n, l: gb.Graph(grp, n, l))
        g.linearize()
        g.optimize(n_iter=3, use_lm=False)
        st = g.optimize(n_iter=20, use_lm=False)
        out = {"variant": spec, "ms_per_iter": st.total_ms / 20, "error_final": st.error_final}
        for k, n in STAGES:
            try:
                out[n] = round(g.time_stage(k, 20), 5)
            except Exception as e:  # stage not available for this configuration
                out[n] = None
        print(json.dumps(out), flush=True)
        for k in sets:
            del os.environ[k]
        del g


if __name__ == "__main__":
    main()
