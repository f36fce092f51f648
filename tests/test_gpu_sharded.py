"""GPU: the trajectory-sharded code path (external separators, k_pack_top -> all-reduce -> k_top_solve -> back-substitution)
exercised on ONE GPU: P shard graphs live on the same device, one Python thread each, with an in-process all-reduce.  The
sharded Gauss-Newton / LM iterates must match the unsharded engine to round-off."""
import threading

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

import gpslam_b200 as gb
from gpslam_b200 import shard, synth


def run_sharded(cfg, world, n_iter, use_lm):
    ar = shard.LocalAllreduce(world, gb.lib())
    shards = []
    for r in range(world):
        sb, _ = synth.build(cfg, lambda grp, N, L, r=r: shard.ShardBuilder(lambda g_, n_, l_: gb.Graph(g_, n_, l_), grp, N, L, r, world), finalize=False)
        sb.g.set_allreduce(ar.make(r))
        shards.append(sb)
    for sb in shards:
        sb.finalize(0)
    stats = [None] * world; errs = [None] * world

    def work(r):
        try:
            stats[r] = shards[r].g.optimize(n_iter=n_iter, use_lm=use_lm)
        except Exception as e:  # noqa
            errs[r] = e
            ar.bar.abort()
    th = [threading.Thread(target=work, args=(r,)) for r in range(world)]
    [t.start() for t in th]; [t.join() for t in th]
    for e in errs:
        if e is not None:
            raise e
    N = cfg.n_states
    P = V = None
    for r, sb in enumerate(shards):
        p, v, l = sb.g.get_values()
        for q, s_ in enumerate(sb.ghosts):  # ghost replicas of remote loop-closure endpoints equal their owner's state bit for bit
            o = shard.owner_of(N, world, s_)
            po_, vo_, _ = shards[o].g.get_values()
            assert np.array_equal(p[sb.n_real + q], po_[s_ - shards[o].lo]) and np.array_equal(v[sb.n_real + q], vo_[s_ - shards[o].lo])
        if P is None:
            P = np.zeros((N, p.shape[1])); V = np.zeros((N, v.shape[1]))
        a, b = shard.owned_range(N, r, world)
        P[a:b] = p[a - sb.lo:b - sb.lo]; V[a:b] = v[a - sb.lo:b - sb.lo]
        if r > 0:  # halo copy is bit-identical to the owner's state
            assert np.array_equal(p[0], P[sb.lo]) and np.array_equal(v[0], V[sb.lo])
    return P, V, l, stats, [sb.g.allreduces() for sb in shards]


@pytest.mark.parametrize("name,n,world", [("C3", 400, 2), ("C3", 333, 3), ("C2", 300, 4), ("C1", 200, 2), ("C4", 500, 3)])
def test_sharded_gn_matches_single(name, n, world):
    cfg = synth.config(name); cfg.n_states = n; cfg.n_landmarks = min(cfg.n_landmarks, 4); cfg.prior_every = 30
    g, _ = synth.build(cfg, lambda grp, N, L: gb.Graph(grp, N, L))
    st = g.optimize(n_iter=3, use_lm=False)
    P0, V0, L0 = g.get_values()
    P, V, Lm, stats, nar = run_sharded(cfg, world, 3, False)
    assert all(k == 3 + 2 for k in nar), nar          # ONE all-reduce per iteration, + one 4-double all-reduce each for the initial and final error report
    assert abs(stats[0].error_final - st.error_final) <= 1e-9 * max(1.0, st.error_final)
    assert np.abs(P - P0).max() < 1e-9 and np.abs(V - V0).max() < 1e-9
    if L0.size:
        assert np.abs(Lm - L0).max() < 1e-9


def test_sharded_lm_matches_single():
    cfg = synth.config("C3"); cfg.n_states = 300; cfg.n_landmarks = 3; cfg.prior_every = 30
    g, _ = synth.build(cfg, lambda grp, N, L: gb.Graph(grp, N, L))
    st = g.optimize(use_lm=True)
    P0, V0, L0 = g.get_values()
    P, V, Lm, stats, _ = run_sharded(cfg, 2, 0, True)
    assert stats[0].iterations == st.iterations
    assert abs(stats[0].error_final - st.error_final) <= 1e-8 * max(1.0, st.error_final)
    assert np.abs(P - P0).max() < 1e-7 and np.abs(V - V0).max() < 1e-7 and np.abs(Lm - L0).max() < 1e-7


@pytest.mark.parametrize("name,n,world,ncl,nl", [("C5", 400, 2, 4, 4), ("C5", 500, 3, 7, 16), ("C1", 240, 2, 5, 4), ("C4", 500, 4, 6, 0)])
def test_sharded_loop_closures_match_single(name, n, world, ncl, nl):
    """loop closures across shards: endpoints join the global reduced system, the evaluating rank carries remote endpoints as ghosts"""
    cfg = synth.config(name); cfg.n_states = n; cfg.n_landmarks = min(cfg.n_landmarks, nl); cfg.prior_every = 30
    cfg.n_closures = ncl; cfg.closure_min_gap = n // 5; cfg.closure_ends = True
    g, _ = synth.build(cfg, lambda grp, N, L: gb.Graph(grp, N, L))
    st = g.optimize(n_iter=3, use_lm=False)
    P0, V0, L0 = g.get_values()
    P, V, Lm, stats, nar = run_sharded(cfg, world, 3, False)
    assert all(k == 3 + 2 for k in nar), nar
    assert abs(stats[0].error_final - st.error_final) <= 1e-9 * max(1.0, st.error_final)
    assert np.abs(P - P0).max() < 1e-9 and np.abs(V - V0).max() < 1e-9
    if L0.size:
        assert np.abs(Lm - L0).max() < 1e-9
    # LM: damping of pinned / ghost entries happens exactly once
    g, _ = synth.build(cfg, lambda grp, N, L: gb.Graph(grp, N, L))
    st = g.optimize(use_lm=True)
    P0, V0, L0 = g.get_values()
    P, V, Lm, stats, _ = run_sharded(cfg, world, 0, True)
    assert stats[0].iterations == st.iterations
    assert abs(stats[0].error_final - st.error_final) <= 1e-8 * max(1.0, st.error_final)
    assert np.abs(P - P0).max() < 1e-7 and np.abs(V - V0).max() < 1e-7
