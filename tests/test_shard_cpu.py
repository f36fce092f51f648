"""CPU, world_size 2 and 3 over gloo: the host-side logic of the trajectory-sharded path (gpslam_b200/shard.py) —
interval ownership (every factor evaluated exactly once), halo / separator bookkeeping and the layout of the all-reduced
boundary system — checked with the oracle standing in for the per-shard engine: each rank eliminates its interior states
densely, the packed Schur systems are summed by torch.distributed (gloo), every rank solves the reduced system and
back-substitutes; the assembled delta must equal the dense solve of the unsharded graph."""
import os
import socket

import numpy as np
import pytest

from gpslam_b200 import shard, synth
from oracle import pyoracle as po


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close(); return p


def _cfg(name, n, ncl):
    cfg = synth.config(name); cfg.n_states = n; cfg.n_landmarks = min(cfg.n_landmarks, 3); cfg.prior_every = 7
    cfg.n_closures = ncl; cfg.closure_min_gap = max(3, n // 6); cfg.closure_ends = ncl >= 2
    return cfg


def _worker(rank, world, port, name, n, out, ncl=0):
    import torch
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"; os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    cfg = _cfg(name, n, ncl)
    sb, _ = synth.build(cfg, lambda grp, N, L: shard.ShardBuilder(lambda g_, n_, l_: po.Graph(g_, n_, l_), grp, N, L, rank, world))
    g = sb.g
    bs, nb = 2 * g.D, g.NL * g.DL
    H, rhs = g.normal_equations_dense()                      # local variables: [local chain entries (own states, ghosts)..., landmarks]
    nloc = g.N
    top = sb.pinned_local                                     # halo / last own state / loop-closure endpoints / ghosts
    top_idx = np.concatenate([np.arange(s * bs, (s + 1) * bs) for s in top] + [np.arange(nloc * bs, nloc * bs + nb)]).astype(int)
    int_idx = np.setdiff1d(np.arange(nloc * bs + nb), top_idx)
    Hii, Hit, Htt = H[np.ix_(int_idx, int_idx)], H[np.ix_(int_idx, top_idx)], H[np.ix_(top_idx, top_idx)]
    S = Htt - Hit.T @ np.linalg.solve(Hii, Hit)
    s = rhs[top_idx] - Hit.T @ np.linalg.solve(Hii, rhs[int_idx])
    R = shard.reduced_dim(world, bs, nb, len(sb.top))
    gi = shard.reduced_index(world, rank, bs, nb, sb.pinned_gtop, len(sb.top))
    T = np.zeros((R, R)); t = np.zeros(R)
    T[np.ix_(gi, gi)] = S; t[gi] = s
    buf = torch.from_numpy(np.concatenate([T.ravel(), t]))
    dist.all_reduce(buf)                                      # the one exchange of the iteration
    buf = buf.numpy()
    x = np.linalg.solve(buf[:R * R].reshape(R, R), buf[R * R:])
    xt = x[gi]
    xi = np.linalg.solve(Hii, rhs[int_idx] - Hit @ xt)
    full = np.zeros(nloc * bs + nb); full[top_idx] = xt; full[int_idx] = xi
    lo, hi = shard.local_range(n, rank, world)
    a, b = shard.owned_range(n, rank, world)
    np.save(os.path.join(out, "delta_%d.npy" % rank), np.concatenate([[a, b, lo], full]))
    dist.destroy_process_group()


@pytest.mark.parametrize("name,world,n,ncl", [("C3", 2, 41, 0), ("C3", 3, 50, 0), ("C1", 2, 37, 0), ("C4", 3, 46, 0),
                                              ("C5", 2, 40, 3), ("C5", 3, 52, 5), ("C1", 3, 45, 4), ("VW", 2, 41, 0), ("VW", 3, 52, 3)])
def test_sharded_schur_matches_full(tmp_path, name, world, n, ncl):
    """ncl > 0: loop closures - endpoints join the reduced system, remote endpoints are carried as ghosts by the evaluating rank"""
    import torch.multiprocessing as mp
    port = _free_port()
    mp.spawn(_worker, args=(world, port, name, n, str(tmp_path), ncl), nprocs=world, join=True)
    cfg = _cfg(name, n, ncl)
    o, _ = synth.build(cfg, lambda grp, N, L: po.Graph(grp, N, L))
    H, rhs = o.normal_equations_dense()
    ref = np.linalg.solve(H, rhs)
    bs, nb = 2 * o.D, o.NL * o.DL
    got = np.zeros_like(ref)
    for r in range(world):
        d = np.load(os.path.join(str(tmp_path), "delta_%d.npy" % r))
        a, b, lo = int(d[0]), int(d[1]), int(d[2]); full = d[3:]
        got[a * bs:b * bs] = full[(a - lo) * bs:(b - lo) * bs]          # owned states
        if r > 0:                                                          # halo copy equals the owner's value
            np.testing.assert_allclose(full[:bs], ref[lo * bs:(lo + 1) * bs], atol=1e-9 * max(1, np.abs(ref).max()))
        got[n * bs:] = full[-nb:] if nb else []
    np.testing.assert_allclose(got, ref, atol=1e-8 * max(1.0, np.abs(ref).max()))


def test_every_factor_owned_once():
    for name, n, ncl in (("C3", 97, 0), ("C1", 64, 0), ("C4", 80, 0), ("C5", 90, 6), ("VW", 75, 4)):
        for world in (2, 3, 5):
            cfg = _cfg(name, n, ncl)
            full, _ = synth.build(cfg, lambda grp, N, L: po.Graph(grp, N, L))
            total = 0; err = 0.0
            for r in range(world):
                sb, _ = synth.build(cfg, lambda grp, N, L: shard.ShardBuilder(lambda g_, n_, l_: po.Graph(g_, n_, l_), grp, N, L, r, world))
                total += sb.g.num_factors(); err += sb.g.error()
            assert total == full.num_factors()
            assert abs(err - full.error()) <= 1e-9 * full.error()


def test_recorder_replays_the_same_graph():
    """synth.record: the generator runs once, its construction calls replay into any number of graph objects (parity tests at
    BASELINE sizes build the engine graph and the oracle graph from one recording)"""
    from gpslam_b200 import synth
    from oracle import pyoracle as po
    cfg = synth.config("C5"); cfg.n_states = 600; cfg.n_landmarks = 4; cfg.n_closures = 3; cfg.closure_min_gap = 100
    rec, truth = synth.record(cfg)
    a = rec.replay(lambda grp, n, l: po.Graph(grp, n, l))
    b, _ = synth.build(cfg, lambda grp, n, l: po.Graph(grp, n, l))
    assert a.num_factors() == b.num_factors() and a.error() == b.error()
    Pa, Va, La = a.get_values(); Pb, Vb, Lb = b.get_values()
    assert np.array_equal(Pa, Pb) and np.array_equal(Va, Vb) and np.array_equal(La, Lb)


def test_oracle_step_check_separates_right_from_wrong():
    """oracle.check_step (the full-size parity instrument of tests/test_gpu_fullsize.py): the exact solution of the oracle's own
    normal equations passes at rounding level, a step that is off by 1e-6 of its size in ONE entry fails by orders of magnitude"""
    from gpslam_b200 import synth
    from oracle import pyoracle as po
    cfg = synth.config("C5"); cfg.n_states = 300; cfg.n_landmarks = 4; cfg.prior_every = 40; cfg.n_closures = 4; cfg.closure_min_gap = 50
    o, _ = synth.build(cfg, lambda grp, n, l: po.Graph(grp, n, l))
    o.set_threads(3)
    o.optimize(n_iter=3, use_lm=True)
    H, g = o.normal_equations_dense()
    for lam in (0.0, 1e-2):
        d = np.linalg.solve(H + lam * np.eye(len(g)), g)
        ns = 300 * 12
        c = o.check_step(d[:ns], d[ns:], lam)
        assert c["residual"] <= 1e-12 * c["scale"], c
        bad = d.copy(); k = int(np.argmax(np.abs(d))); bad[k] *= 1 + 1e-6
        cb = o.check_step(bad[:ns], bad[ns:], lam)
        assert cb["residual"] > 1e-8 * cb["scale"] and cb["residual"] > 1e3 * c["residual"]
