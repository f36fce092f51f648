"""Small end-to-end runs of every solver path, meant to be executed under compute-sanitizer (scripts/sanitize.sh):
SE(3) with a 64-column panel (k_spine + k_panel4, k_level_ws, k_small_solve), with loop closures (pinned separators, blocked dense
top solve), narrow borders (generic k_fwd / k_bwd), SO(3) / SE(2) chains, a 2-shard graph (pack -> all-reduce -> redundant solve)
and the pipelined batch interface.  Results are compared with the oracle, so a sanitizer-clean run is also a correct one."""
import os
import sys
import threading

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import gpslam_b200 as gb  # noqa: E402
from gpslam_b200 import shard, synth  # noqa: E402
from oracle import pyoracle as po  # noqa: E402

which = sys.argv[1] if len(sys.argv) > 1 else "all"
n_iter = int(sys.argv[2]) if len(sys.argv) > 2 else 2


def cfg_of(name, n, **kw):
    c = synth.config(name); c.n_states = n
    for k, v in kw.items():
        setattr(c, k, v)
    return c


CASES = {
    "pose3_wide": cfg_of("C3", 333, n_landmarks=16, prior_every=40),
    "pose3_wide_loops": cfg_of("C5", 500, n_landmarks=16, prior_every=40, n_closures=8, closure_min_gap=60, closure_ends=True),
    "pose3": cfg_of("C3", 300, n_landmarks=4, prior_every=40),
    "pose2": cfg_of("C1", 200),
    "rot3": cfg_of("C4", 301),
    "many_closures": cfg_of("C5", 1500, n_landmarks=4, prior_every=100, n_closures=20, closure_min_gap=150),
}


def check(tag, g, o):
    Pg, Vg, Lg = g.get_values(); Po, Vo, Lo = o.get_values()
    err = max(np.abs(Pg - Po).max(), np.abs(Vg - Vo).max(), np.abs(Lg - Lo).max() if Lo.size else 0.0)
    print("%-22s max |gpu - cpu| = %.2e" % (tag, err), flush=True)
    assert err < 1e-6, (tag, err)


for name, cfg in CASES.items():
    if which not in ("all", name):
        continue
    rec, _ = synth.record(cfg)
    for use_lm in (False, True):
        g = rec.replay(lambda grp, n, l: gb.Graph(grp, n, l)); o = rec.replay(lambda grp, n, l: po.Graph(grp, n, l))
        g.optimize(n_iter=n_iter, use_lm=use_lm); o.optimize(n_iter=n_iter, use_lm=use_lm)
        check("%s %s" % (name, "LM" if use_lm else "GN"), g, o)

if which in ("all", "batch"):
    cfg = CASES["pose3_wide"]
    rec, _ = synth.record(cfg)
    g = rec.replay(lambda grp, n, l: gb.Graph(grp, n, l)); o = rec.replay(lambda grp, n, l: po.Graph(grp, n, l))
    ins = [g.get_values(out=g.alloc_values()) for _ in range(2)]; outs = [g.alloc_values() for _ in range(2)]
    g.optimize_batch([ins[k & 1] for k in range(4)], [outs[k & 1] for k in range(4)])
    o.optimize(n_iter=1, use_lm=False)
    Po, Vo, Lo = o.get_values()
    for k in range(2):
        err = max(np.abs(outs[k][0] - Po).max(), np.abs(outs[k][1] - Vo).max(), np.abs(outs[k][2][:, :3] - Lo).max())
        print("batch out[%d]           max |gpu - cpu| = %.2e" % (k, err), flush=True)
        assert err < 1e-6

if which in ("all", "sharded"):
    cfg = cfg_of("C5", 500, n_landmarks=16, prior_every=30, n_closures=5, closure_min_gap=100, closure_ends=True)
    world = 2
    ar = shard.LocalAllreduce(world, gb.lib())
    shards = []
    for r in range(world):
        sb, _ = synth.build(cfg, lambda grp, N, L, r=r: shard.ShardBuilder(lambda g_, n_, l_: gb.Graph(g_, n_, l_), grp, N, L, r, world), finalize=False)
        sb.g.set_allreduce(ar.make(r)); shards.append(sb)
    for sb in shards:
        sb.finalize(0)
    th = [threading.Thread(target=lambda r=r: shards[r].g.optimize(n_iter=n_iter, use_lm=False)) for r in range(world)]
    [t.start() for t in th]; [t.join() for t in th]
    o, _ = synth.build(cfg, lambda grp, n, l: po.Graph(grp, n, l))
    o.optimize(n_iter=n_iter, use_lm=False)
    Po, Vo, Lo = o.get_values()
    for r, sb in enumerate(shards):
        p, v, l = sb.g.get_values()
        a, b = shard.owned_range(cfg.n_states, r, world)
        err = max(np.abs(p[a - sb.lo:b - sb.lo] - Po[a:b]).max(), np.abs(v[a - sb.lo:b - sb.lo] - Vo[a:b]).max(), np.abs(l - Lo).max())
        print("sharded rank %d         max |gpu - cpu| = %.2e" % (r, err), flush=True)
        assert err < 1e-6
print("sanitize_cases ok", flush=True)
