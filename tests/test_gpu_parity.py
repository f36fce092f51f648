"""GPU parity tests: the CUDA engine (through the C ABI, gpslam_b200/libgpb.so) against the CPU oracle on identical seeded
graphs.  Tolerances: whitened residuals/rhs 1e-9 relative; Jacobian blocks 1e-6 (the oracle keeps the reference's 1e-6-step
numerical differentiation of rightJacobianPose3inv, the CUDA path differentiates in closed form); block-solver vs dense
solve of its own normal equations 1e-8 relative; optimised states <= 1e-6 (north_star tolerance)."""
import math

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

import gpslam_b200 as gb
from gpslam_b200 import synth
from oracle import pyoracle as po

POSE3, POSE2, ROT3, LINEAR = 0, 1, 2, 3


def both(cfg, seglen=None):
    def mk(grp, n, l):
        g = gb.Graph(grp, n, l)
        if seglen:
            g.set_segment_length(*seglen)
        return g
    g, truth = synth.build(cfg, mk)
    o, _ = synth.build(cfg, lambda grp, n, l: po.Graph(grp, n, l))
    return g, o, truth


def small_cfg(name, n, **kw):
    cfg = synth.config(name); cfg.n_states = n
    for k, v in kw.items():
        setattr(cfg, k, v)
    return cfg


CASES = {
    "pose3": dict(name="C3", n=300, n_landmarks=4, prior_every=40),
    # 16 landmarks -> 64-column panel: the production SE(3) kernels (k_spine + k_panel4), which narrower borders never reach
    "pose3_wide": dict(name="C3", n=333, n_landmarks=16, prior_every=40),
    "pose3_chain": dict(name="C2", n=257, prior_every=30),
    "pose2": dict(name="C1", n=200),
    "rot3": dict(name="C4", n=301),
    # loop closures (BetweenFactor between distant states, BASELINE config C5): endpoint states become pinned separators and
    # the Schur complement on {endpoints, landmarks} is solved densely
    # GPS fixes and pinhole projections through the GP interpolator, incl. one landmark behind its camera
    "pose3_gps_proj": dict(name="C3", n=260, n_landmarks=6, prior_every=40, gps_every=7, proj_per_state=0.3),
    # SE(3) "VW" family: GaussianProcessPriorPose3VW + GPInterpolatedGPSFactorPose3VW, velocities
    # [v_world | w_world]; with and without loop closures (no landmarks: the generic k_fwd<12,16> solver path)
    "pose3vw": dict(name="VW", n=280, prior_every=40, gps_every=3),
    "pose3vw_loops": dict(name="VW", n=300, prior_every=50, gps_every=4, n_closures=4, closure_min_gap=40),
    # a correlated (dense) Qc: the dense-Rq instantiation of k_lin_gp (diagonal Qc models take the element-wise one)
    "pose3_dense_qc": dict(name="C3", n=200, n_landmarks=4, prior_every=40, qc_dense=True),
    "pose2_dense_qc": dict(name="C1", n=150, qc_dense=True),
    "pose3_loops": dict(name="C5", n=400, n_landmarks=4, prior_every=40, n_closures=5, closure_min_gap=40),
    "pose3_wide_loops": dict(name="C5", n=500, n_landmarks=16, prior_every=40, n_closures=8, closure_min_gap=60, closure_ends=True),
    "pose2_loops": dict(name="C1", n=200, n_closures=4, closure_min_gap=30, closure_ends=True),
    "rot3_loops": dict(name="C4", n=301, n_closures=3, closure_min_gap=50),
}


def linear_graph(make, n=120, seed=5):
    """Linear<3> ("2DLinear") trajectory with every plain 2-D factor of gpslam/slam"""
    rng = np.random.default_rng(seed)
    cfg = small_cfg("C1", n); cfg.group = LINEAR; cfg.odometry = False
    poses, vels = synth.ground_truth(cfg)
    L = 3
    g = make(LINEAR, n, L)
    g.add_qc_model(np.eye(3) * 0.01)
    g.add_gp_prior(np.arange(n - 1), np.full(n - 1, cfg.dt))
    lands = rng.uniform(-20, 20, size=(L, 2)) + poses[:, :2].mean(axis=0)
    ri = np.sort(rng.integers(0, n - 1, size=n // 2)); rl = rng.integers(0, L, size=len(ri)); tau = rng.uniform(-0.02, cfg.dt + 0.02, size=len(ri))
    z = np.array([np.linalg.norm(lands[l] - (poses[i, :2] + vels[i, :2] * t)) for i, l, t in zip(ri, rl, tau)]) + rng.normal(size=len(ri)) * 0.3
    g.add_interp_range(ri, rl, z, np.full(len(ri), 0.5), np.full(len(ri), cfg.dt), tau)
    for i in range(0, n, 7):
        l = int(rng.integers(0, L)); d = lands[l] - poses[i, :2]
        g.add_range_2d(i, l, float(np.linalg.norm(d) + rng.normal() * 0.2), 0.4)
        c, s = np.cos(poses[i, 2]), np.sin(poses[i, 2])
        g.add_range_bearing_2d(i, (l + 1) % L, float(np.linalg.norm(lands[(l + 1) % L] - poses[i, :2])),
                               float(math.atan2(-s * (lands[(l + 1) % L] - poses[i, :2])[0] + c * (lands[(l + 1) % L] - poses[i, :2])[1],
                                                c * (lands[(l + 1) % L] - poses[i, :2])[0] + s * (lands[(l + 1) % L] - poses[i, :2])[1]) + rng.normal() * 0.01),
                               np.array([[20.0, 1.0], [0.0, 3.0]]))
    for i in range(n - 1):
        c, s = np.cos(poses[i, 2]), np.sin(poses[i, 2]); d = poses[i + 1] - poses[i]
        g.add_odometry_2d(i, i + 1, np.array([c * d[0] + s * d[1], -s * d[0] + c * d[1], d[2]]) + rng.normal(size=3) * 0.01, np.diag([50.0, 50.0, 100.0]))
    for l in range(L):
        g.add_prior_landmark(l, lands[l] + rng.normal(size=2) * 0.3, np.eye(2))
    g.add_prior_pose(0, poses[0], np.eye(3) * 10); g.add_prior_vel(0, vels[0], np.eye(3) * 10)
    g.add_prior_pose(n - 1, poses[n - 1] + 0.05, np.array([[5.0, 0.5, 0.1], [0, 4.0, 0.2], [0, 0, 3.0]]))
    g.add_prior_vel(n - 1, vels[n - 1], np.eye(3) * 2)
    init = poses + rng.normal(size=poses.shape) * 0.05
    g.set_values(init, np.zeros((n, 3)), lands + rng.normal(size=lands.shape) * 0.3)
    if hasattr(g, "finalize"):
        g.finalize()
    return g


def pose2_range2d_graph(make, n=160, seed=9):
    """Pose2 trajectory (Plaza shape) whose range measurements are plain RangeFactorPose2 (slam/RangeFactorPose2.h:15 =
    gtsam::RangeFactor<Pose2, Point2>) at the states themselves, beside interpolated ranges between them"""
    rng = np.random.default_rng(seed)
    cfg = small_cfg("C1", n)
    poses, vels = synth.ground_truth(cfg)
    L = 4
    g = make(POSE2, n, L)
    g.add_qc_model(np.eye(3) * cfg.qc_sigma ** 2)
    g.add_gp_prior(np.arange(n - 1), np.full(n - 1, cfg.dt))
    lands = poses[:, :2].mean(axis=0) + rng.uniform(-25, 25, size=(L, 2))
    for i in list(range(0, n, 3)) + [n - 1]:   # incl. the last state (the b-part of the last interval)
        l = int(rng.integers(0, L))
        g.add_range_2d(i, l, float(np.linalg.norm(lands[l] - poses[i, :2]) + rng.normal() * 0.3), 0.5)
    ri = np.sort(rng.integers(0, n - 1, size=n // 4)); rl = rng.integers(0, L, size=len(ri)); tau = rng.uniform(0, cfg.dt, size=len(ri))
    z = np.array([np.linalg.norm(lands[l] - synth._retract(POSE2, poses[i], vels[i] * t)[:2]) for i, l, t in zip(ri, rl, tau)]) + rng.normal(size=len(ri)) * 0.3
    g.add_interp_range(ri, rl, z, np.full(len(ri), 0.5), np.full(len(ri), cfg.dt), tau)
    for i in range(n - 1):
        g.add_between(i, i + 1, synth._between(POSE2, poses[i], poses[i + 1], rng, 1e-3), np.diag([1e3, 1e3, 1e3 / np.pi]))
    for l in range(L):
        g.add_prior_landmark(l, lands[l] + rng.normal(size=2) * 0.5, np.eye(2))
    g.add_prior_pose(0, poses[0], np.eye(3)); g.add_prior_vel(0, vels[0], np.eye(3))
    g.set_values(np.stack([synth._retract(POSE2, poses[i], rng.normal(size=3) * 0.05) for i in range(n)]), np.zeros((n, 3)), lands + rng.normal(size=lands.shape) * 0.5)
    if hasattr(g, "finalize"):
        g.finalize()
    return g


def make_pair(case):
    if case == "pose2_range2d":
        return pose2_range2d_graph(lambda grp, n, l: gb.Graph(grp, n, l)), pose2_range2d_graph(lambda grp, n, l: po.Graph(grp, n, l))
    if case == "linear":
        return linear_graph(lambda grp, n, l: gb.Graph(grp, n, l)), linear_graph(lambda grp, n, l: po.Graph(grp, n, l))
    kw = dict(CASES[case]); name = kw.pop("name"); n = kw.pop("n")
    g, o, _ = both(small_cfg(name, n, **kw))
    return g, o


ALL = ["pose3", "pose3_wide", "pose3_chain", "pose2", "pose2_range2d", "rot3", "linear", "pose3_gps_proj", "pose3vw", "pose3vw_loops", "pose3_dense_qc", "pose2_dense_qc", "pose3_loops", "pose3_wide_loops", "pose2_loops", "rot3_loops"]


@pytest.mark.parametrize("case", ALL)
def test_linearize_matches_oracle(case):
    g, o = make_pair(case)
    e_gpu = g.linearize(); e_cpu = o.error()
    assert abs(e_gpu - e_cpu) <= 1e-9 * max(1.0, abs(e_cpu))
    assert abs(g.error() - e_cpu) <= 1e-9 * max(1.0, abs(e_cpu))  # residual-only path
    tolA = 1e-6 if case.startswith("pose3") else 1e-9
    nint = g.N - 1
    nf = o.num_factors()
    # the generators add the GP priors first (factor k = interval k), then the other factors in insertion order
    for k in list(range(0, nint, max(1, nint // 60))) + [nint - 1]:
        Ao, bo = o.linearize_factor(k); Ag, bg = g.linearized_factor(0, k)
        sc = max(1.0, max(np.abs(a).max() for a in Ao))
        np.testing.assert_allclose(bg, bo, atol=1e-9 * max(1.0, np.abs(bo).max()))
        for x, y in zip(Ag, Ao):
            np.testing.assert_allclose(x, y, atol=tolA * sc)
    for idx in range(nf - nint):
        Ao, bo = o.linearize_factor(nint + idx); Ag, bg = g.linearized_factor(1, idx)
        assert len(Ag) == len(Ao)
        sc = max(1.0, max(np.abs(a).max() for a in Ao))
        np.testing.assert_allclose(bg, bo, atol=1e-9 * max(1.0, np.abs(bo).max()))
        for x, y in zip(Ag, Ao):
            np.testing.assert_allclose(x, y, atol=tolA * sc)


@pytest.mark.parametrize("case", ALL)
def test_normal_equations_and_block_solver(case):
    g, o = make_pair(case)
    g.linearize()
    Hg, gg = g.normal_equations_dense()
    Ho, go = o.normal_equations_dense()
    np.testing.assert_allclose(Hg, Hg.T, atol=1e-9 * np.abs(Hg).max())
    np.testing.assert_allclose(Hg, Ho, atol=2e-6 * np.abs(Ho).max())
    np.testing.assert_allclose(gg, go, atol=2e-6 * np.abs(go).max())
    for lam in (0.0, 1e-3, 10.0):
        ds, dl = g.solve_delta(lam)
        x = np.concatenate([ds.ravel(), dl])
        ref = np.linalg.solve(Hg + lam * np.eye(len(gg)), gg)
        assert np.abs(x - ref).max() <= 1e-8 * max(1.0, np.abs(ref).max()), (case, lam, np.abs(x - ref).max())


@pytest.mark.parametrize("case", ALL)
@pytest.mark.parametrize("use_lm", [False, True])
def test_optimize_matches_oracle(case, use_lm):
    g, o = make_pair(case)
    sg = g.optimize(use_lm=use_lm); so = o.optimize(use_lm=use_lm)
    assert sg.status == 0 and so.status == 0
    assert sg.iterations == so.iterations
    assert abs(sg.error_final - so.error_final) <= 1e-7 * max(1.0, so.error_final)
    Pg, Vg, Lg = g.get_values(); Po, Vo, Lo = o.get_values()
    assert np.abs(Pg - Po).max() <= 1e-6
    assert np.abs(Vg - Vo).max() <= 1e-6
    if Lo.size:
        assert np.abs(Lg - Lo).max() <= 1e-6


@pytest.mark.parametrize("variant", [0, 1, 2, 3])
def test_lin_gp_instantiations_agree(variant, monkeypatch):
    """k_lin_gp<G_POSE3> instantiations (GPB_LIN_VARIANT: dense / diagonal Rq, two / three CTAs per SM) produce the same [A|b]"""
    cfg = small_cfg("C3", 300, n_landmarks=4, prior_every=40)
    g0, o, _ = both(cfg)
    e0 = g0.linearize()
    monkeypatch.setenv("GPB_LIN_VARIANT", str(variant))
    g1, _ = synth.build(cfg, lambda grp, n, l: gb.Graph(grp, n, l))
    e1 = g1.linearize()
    assert abs(e1 - e0) <= 1e-12 * abs(e0) and abs(e1 - o.error()) <= 1e-9 * abs(e0)
    for k in range(0, g0.N - 1, 7):
        A0, b0 = g0.linearized_factor(0, k); A1, b1 = g1.linearized_factor(0, k)
        sc = max(1.0, max(np.abs(a).max() for a in A0))
        np.testing.assert_allclose(b1, b0, atol=1e-12 * max(1.0, np.abs(b0).max()))
        for x, y in zip(A1, A0):
            np.testing.assert_allclose(x, y, atol=1e-12 * sc)


@pytest.mark.parametrize("n", [5, 8, 9, 127, 128, 129, 130, 300])
def test_assembly_kernels_agree(n, monkeypatch):
    """SE(3) assembly: the tensor-pipe kernel (k_assemble_mma, default) against the thread-per-tile kernel (GPB_OLD_ASSEMBLE) on
    chains that end inside / at / just behind a tile of the [A|b] layout (128 factors) and of the kernel (8 states)"""
    cfg = small_cfg("C3", n, n_landmarks=4, prior_every=max(2, min(40, n // 2)))
    g0, _ = synth.build(cfg, lambda grp, nn, l: gb.Graph(grp, nn, l))
    monkeypatch.setenv("GPB_OLD_ASSEMBLE", "1")
    g1, _ = synth.build(cfg, lambda grp, nn, l: gb.Graph(grp, nn, l))
    g0.linearize(); g1.linearize()
    H0, r0 = g0.normal_equations_dense(); H1, r1 = g1.normal_equations_dense()
    np.testing.assert_allclose(H0, H1, atol=1e-12 * np.abs(H1).max()); np.testing.assert_allclose(r0, r1, atol=1e-12 * np.abs(r1).max())


@pytest.mark.parametrize("n_landmarks", [2, 16])
@pytest.mark.parametrize("n", [2, 3, 16, 17, 18, 33, 129, 130])
def test_segment_boundaries(n, n_landmarks):
    """chain lengths around the segment cuts, several segment lengths, narrow (generic kernel) and 64-column (spine + panel
    kernels) borders: same delta as a dense solve"""
    for seglen in ((2, 2), (4, 3), (16, 8), None):
        cfg = small_cfg("C3", n, n_landmarks=n_landmarks, prior_every=5, range_per_state=0.7)
        g, o, _ = both(cfg, seglen)
        g.linearize()
        Hg, gg = g.normal_equations_dense()
        ds, dl = g.solve_delta(0.0)
        ref = np.linalg.solve(Hg, gg)
        x = np.concatenate([ds.ravel(), dl])
        assert np.abs(x - ref).max() <= 1e-8 * max(1.0, np.abs(ref).max()), (n, seglen)


@pytest.mark.parametrize("seglen", [None, (7, 3)])
def test_many_loop_closures_blocked_top_solve(seglen):
    """C5 shape with enough closures that the reduced system (40 endpoint states x 12 + landmarks = 492 unknowns) goes through
    the blocked multi-CTA Cholesky: one damped step against a solve of the same system by the oracle's bordered solver is not
    available as a dense matrix at this size, so compare the Gauss-Newton / LM trajectories instead."""
    cfg = small_cfg("C5", 3000, n_landmarks=4, prior_every=100, n_closures=20, closure_min_gap=300)
    g, o, _ = both(cfg, seglen)
    assert g.sizes().levels >= 2
    sg = g.optimize(n_iter=1, use_lm=False); so = o.optimize(n_iter=1, use_lm=False)
    assert abs(sg.error_final - so.error_final) <= 1e-7 * max(1.0, so.error_final)
    Pg, Vg, Lg = g.get_values(); Po, Vo, Lo = o.get_values()
    assert np.abs(Pg - Po).max() <= 1e-6 and np.abs(Vg - Vo).max() <= 1e-6 and np.abs(Lg - Lo).max() <= 1e-6
    sg = g.optimize(use_lm=True); so = o.optimize(use_lm=True)
    assert sg.status == 0 and sg.iterations == so.iterations
    assert abs(sg.error_final - so.error_final) <= 1e-7 * max(1.0, so.error_final)
    Pg, Vg, Lg = g.get_values(); Po, Vo, Lo = o.get_values()
    assert np.abs(Pg - Po).max() <= 1e-6 and np.abs(Vg - Vo).max() <= 1e-6 and np.abs(Lg - Lo).max() <= 1e-6


@pytest.mark.parametrize("R", [1, 2, 15, 16, 17, 31, 32, 47, 48, 49, 63, 64, 65, 95, 96, 128, 132, 143, 144, 160, 161, 200, 449, 1000])
def test_reduced_system_solvers(R):
    """the reduced-system solvers alone (shared-memory single-CTA solver in its register-blocked (R <= 143) and plain-loop
    instantiations, blocked multi-CTA Cholesky) against numpy"""
    from gpslam_b200 import capi
    rng = np.random.default_rng(R)
    B = rng.normal(size=(R, R + 5)); A = B @ B.T + 0.1 * np.eye(R); b = rng.normal(size=R)
    for lam, loff in ((0.0, 0), (0.7, R // 2)):
        ref = np.linalg.solve(A + lam * np.diag((np.arange(R) >= loff).astype(float)), b)
        for blocked in ((False, True) if R <= 160 else (True,)):
            x = capi.dense_solve(A, b, lam, loff, blocked)
            assert np.abs(x - ref).max() <= 1e-9 * max(1.0, np.abs(ref).max()), (R, lam, blocked, np.abs(x - ref).max())
        if R <= 143:
            for kw in (dict(force_small=True), dict(old_tiny=True)):
                x = capi.dense_solve(A, b, lam, loff, **kw)
                assert np.abs(x - ref).max() <= 1e-9 * max(1.0, np.abs(ref).max()), (R, lam, kw, np.abs(x - ref).max())
    with pytest.raises(capi.GpbError):
        capi.dense_solve(-A, b, 0.0, 0, R > 160)
    if R <= 143:
        with pytest.raises(capi.GpbError):
            capi.dense_solve(-A, b, 0.0, 0, force_small=True)


def test_dense_solve_trailing_update_variants(monkeypatch):
    """the multi-CTA dense solve with its trailing update on the tensor pipe (default) and on FP64 FMAs (GPB_FMA_SYRK) against numpy"""
    from gpslam_b200 import capi
    rng = np.random.default_rng(77)
    for R in (130, 449, 1000):
        B = rng.normal(size=(R, R + 5)); A = B @ B.T + 0.1 * np.eye(R); b = rng.normal(size=R)
        ref = np.linalg.solve(A + 0.3 * np.diag((np.arange(R) >= R // 3).astype(float)), b)
        x_mma = capi.dense_solve(A, b, 0.3, R // 3, True)
        monkeypatch.setenv("GPB_FMA_SYRK", "1")
        x_fma = capi.dense_solve(A, b, 0.3, R // 3, True)
        monkeypatch.delenv("GPB_FMA_SYRK")
        for x in (x_mma, x_fma):
            assert np.abs(x - ref).max() <= 1e-9 * max(1.0, np.abs(ref).max()), (R, np.abs(x - ref).max())


@pytest.mark.parametrize("case", ["pose3_wide", "pose2", "rot3"])
def test_optimize_batch_matches_oracle(case):
    """gpb_optimize_batch: K independent (values in -> one GN iteration -> values out) steps, copies pipelined against compute.  Two
    different inputs alternate; every step's output and error equal the oracle's single iteration from that input"""
    g, o = make_pair(case)
    P0, V0, L0 = g.get_values()
    inA = g.get_values(out=g.alloc_values())
    inB = g.alloc_values()
    rng = np.random.default_rng(3)
    inB[0][:] = P0; inB[1][:] = V0 + rng.normal(size=V0.shape) * 0.01
    if L0.size:
        inB[2][:, :L0.shape[1]] = L0 + 0.05
    outs = [g.alloc_values() for _ in range(4)]
    st, errs = g.optimize_batch([inA, inB, inA, inB], outs)
    assert st.status == 0 and st.iterations == 4
    for k, (pin, vin, lin) in enumerate([inA, inB]):
        o.set_values(pin, vin, lin[:, :L0.shape[1]] if L0.size else None)
        so = o.optimize(n_iter=1, use_lm=False)
        Po, Vo, Lo = o.get_values()
        for kk in (k, k + 2):
            assert np.abs(outs[kk][0] - Po).max() <= 1e-6 and np.abs(outs[kk][1] - Vo).max() <= 1e-6
            if L0.size:
                assert np.abs(outs[kk][2][:, :L0.shape[1]] - Lo).max() <= 1e-6
            assert abs(errs[kk] - so.error_final) <= 1e-7 * max(1.0, so.error_final)
        assert np.array_equal(outs[k][0], outs[k + 2][0]) and errs[k] == errs[k + 2]   # same input, same bits
    # the graph is left at the last step's result and keeps working through the ordinary calls
    Pn, Vn, Ln = g.get_values()
    assert np.array_equal(Pn, outs[3][0])
    assert abs(g.linearize() - errs[3]) <= 1e-9 * max(1.0, errs[3])


def test_reference_two_state_optimizations():
    """the reference's 'Optimization' unit tests through the CUDA path (gp/tests/testGaussianProcessPriorPose3.cpp:146-195,
    slam/tests/testGPInterpolatedRangeFactorPose3.cpp:177-260 incl. extrapolation tau = -0.1 and 0.2)"""
    from tests.test_oracle_golden import P3, iso, _range3
    p1, p2 = P3(0, 0, 0, 0, 0, 0), P3(0, 0, 0, 1, 0, 0)
    g = gb.Graph(POSE3, 2, 0)
    g.add_qc_model(0.01 * np.eye(6)); g.add_prior_pose(0, p1, iso(6, 0.001)); g.add_prior_pose(1, p2, iso(6, 0.001)); g.add_gp_prior(0, 1.0)
    g.set_values(np.stack([p1, p2]), np.array([[0, 0, 0, 1, 0, 0], [.1, .2, -.3, 2., -.5, .6]]))
    g.finalize()
    st = g.optimize(use_lm=False)
    P, V, _ = g.get_values()
    assert st.error_final < 1e-6
    np.testing.assert_allclose(P, [p1, p2], atol=1e-6); np.testing.assert_allclose(V, [[0, 0, 0, 1, 0, 0]] * 2, atol=1e-6)

    # slam/tests/testGPInterpolatedGPSFactorPose3.cpp:179-255 and testGPInterpolatedProjectionFactorPose3.cpp:191-268
    from tests.test_oracle_golden import _project
    v = [0, 0, 0, 10, 0, 0]
    g = gb.Graph(POSE3, 2, 0)
    g.add_qc_model(0.01 * np.eye(6)); g.add_prior_pose(0, p1, iso(6, 100.0)); g.add_prior_vel(0, v, iso(6, 0.01)); g.add_prior_vel(1, v, iso(6, 0.01)); g.add_gp_prior(0, 0.1)
    for x, tau in zip((-1, .5, 2), (-.1, .05, .2)):
        g.add_interp_gps(0, [x, 0, 0], iso(3, 0.1), 0.1, tau)
    g.set_values(np.stack([P3(.1, .1, -.1, .04, .1, -.06), P3(-.1, .1, -.1, 1.05, -.1, .1)]), np.array([[-.1, 0, 0, 9.8, 0, .2], [0, 0, .2, 9.7, 0, -.1]]))
    g.finalize()
    st = g.optimize(use_lm=False)
    P, V, _ = g.get_values()
    assert st.error_final < 1e-6
    np.testing.assert_allclose(P, [p1, p2], atol=1e-6); np.testing.assert_allclose(V, [v, v], atol=1e-6)
    K = [50, 50, 0, 40, 30]; land = [3.4, 1.2, 20]
    g = gb.Graph(POSE3, 2, 1)
    g.add_qc_model(0.01 * np.eye(6)); g.add_prior_pose(0, p1, iso(6, 0.01)); g.add_prior_pose(1, p2, iso(6, 0.01)); g.add_gp_prior(0, 0.1)
    for x, tau in zip((.2, .6, .9), (.02, .06, .09)):
        g.add_interp_projection(0, 0, _project(P3(0, 0, 0, x, 0, 0), K, land), iso(2, 0.1), 0.1, tau, K)
    g.set_values(np.stack([P3(.1, .2, .4, .2, .3, -.2), P3(-.1, -.2, -.4, 1.2, -.3, .2)]), np.array([[-.3, 0, 0, .7, 0, .2], [0, 0, .4, 1.2, 0, -.1]]), np.array([[3.3, 1.3, 18]]))
    g.finalize()
    st = g.optimize(use_lm=False)
    P, V, Lm = g.get_values()
    assert st.error_final < 1e-6
    np.testing.assert_allclose(P, [p1, p2], atol=1e-6); np.testing.assert_allclose(V, [v, v], atol=1e-6); np.testing.assert_allclose(Lm[0], land, atol=1e-6)

    land = np.array([.4, 1.2, 3.0]); v = [0, 0, 0, 10, 0, 0]
    meas = [_range3(P3(0, 0, 0, x, 0, 0), land) for x in (-1, .5, 2)]
    g = gb.Graph(POSE3, 2, 1)
    g.add_qc_model(0.01 * np.eye(6))
    g.add_prior_pose(0, p1, iso(6, 0.01)); g.add_prior_pose(1, p2, iso(6, 0.01)); g.add_prior_landmark(0, land, iso(3, 0.1))
    g.add_prior_vel(0, v, iso(6, 0.01)); g.add_prior_vel(1, v, iso(6, 0.01)); g.add_gp_prior(0, 0.1)
    for m, tau in zip(meas, (-.1, .05, .2)):
        g.add_interp_range(0, 0, m, 0.1, 0.1, tau)
    g.set_values(np.stack([P3(.1, .2, .4, .2, .3, -.2), P3(-.1, -.2, -.4, 1.2, -.3, .2)]), np.array([[-.1, 0, 0, .8, 0, .2], [0, 0, .2, 1.2, 0, -.1]]),
                 np.array([[.3, 1.1, 2.9]]))
    g.finalize()
    st = g.optimize(use_lm=False)
    P, V, Lm = g.get_values()
    assert st.error_final < 1e-6
    np.testing.assert_allclose(P, [p1, p2], atol=1e-6); np.testing.assert_allclose(V, [v, v], atol=1e-6); np.testing.assert_allclose(Lm[0], land, atol=1e-6)


def test_full_size_properties():
    """BASELINE config C3 at full size (100k SE(3) states, 50k interpolated ranges, 16 landmarks), size-independent properties:
    (i) two engine instances with different segment lengths (= different elimination orders of the same normal equations)
    return the same Gauss-Newton / damped step; (ii) LM never increases the error (every accepted step passed the fidelity
    test); (iii) both instances converge (GTSAM's stop rule) to the same solution to 1e-6; (iv) one Gauss-Newton iteration
    from the initial values lands where the CPU oracle's does (<= 1e-6, north_star's tolerance).
    Not asserted: that the GN step vanishes at the LM optimum - the tail of the trajectory (range-only beyond the last pose
    prior) holds a weakly observable mode on which plain GN converges linearly, identically in the oracle and in the engine."""
    cfg = synth.config("C3")
    g, truth = synth.build(cfg, lambda grp, n, l: gb.Graph(grp, n, l))
    def mk(grp, n, l):
        h = gb.Graph(grp, n, l); h.set_segment_length(20, 5); return h
    g2, _ = synth.build(cfg, mk)
    e1, e2 = g.linearize(), g2.linearize()
    assert e1 == e2
    for lam in (0.0, 1e-3):
        ds1, dl1 = g.solve_delta(lam); ds2, dl2 = g2.solve_delta(lam)
        assert np.abs(ds1 - ds2).max() <= 1e-8 * max(1.0, np.abs(ds1).max()), (lam, np.abs(ds1 - ds2).max())
        assert np.abs(dl1 - dl2).max() <= 1e-8 * max(1.0, np.abs(dl1).max())
    # (iv) first: one GN iteration of a third instance against the oracle at full size
    g3, _ = synth.build(cfg, lambda grp, n, l: gb.Graph(grp, n, l))
    o, _ = synth.build(cfg, lambda grp, n, l: po.Graph(grp, n, l))
    o.set_threads(po.hardware_threads())
    sg = g3.optimize(n_iter=1, use_lm=False); so = o.optimize(n_iter=1, use_lm=False)
    assert abs(sg.error_final - so.error_final) <= 1e-7 * so.error_final
    Pg, Vg, Lg = g3.get_values(); Po, Vo, Lo = o.get_values()
    assert np.abs(Pg - Po).max() <= 1e-6 and np.abs(Vg - Vo).max() <= 1e-6 and np.abs(Lg - Lo).max() <= 1e-6
    del g3, o
    errs = [e1]
    for _ in range(8):
        st = g.optimize(n_iter=1, use_lm=True); g2.optimize(n_iter=1, use_lm=True)
        errs.append(st.error_final)
    assert all(b <= a * (1 + 1e-12) for a, b in zip(errs, errs[1:])), errs
    st = g.optimize(use_lm=True); st2 = g2.optimize(use_lm=True)
    assert st.status == 0 and st2.status == 0 and st.iterations == st2.iterations
    assert abs(st.error_final - st2.error_final) <= 1e-9 * st.error_final
    P1, V1, L1 = g.get_values(); P2, V2, L2 = g2.get_values()
    assert np.abs(P1 - P2).max() < 1e-6 and np.abs(V1 - V2).max() < 1e-6 and np.abs(L1 - L2).max() < 1e-6


# ----------------------------------------------------------------------------- edge cases
def _ragged_graph(make, n=90, seed=11):
    """SE(3) chain with holes: every 7th interval has NO GP prior (odometry BetweenFactors keep the chain connected), intervals
    with several measurements and intervals with none, one landmark nobody observes (prior only), tau outside [0, dt]"""
    rng = np.random.default_rng(seed)
    cfg = small_cfg("C3", n, n_landmarks=3)
    poses, vels = synth.ground_truth(cfg)
    g = make(POSE3, n, 3)
    g.add_qc_model(np.eye(6) * 0.01)
    keep = np.array([i for i in range(n - 1) if i % 7 != 3])
    g.add_gp_prior(keep, np.full(len(keep), cfg.dt))
    for i in range(n - 1):
        if i % 7 == 3 or i % 5 == 0:
            g.add_between(i, i + 1, synth._between(POSE3, poses[i], poses[i + 1], rng, 1e-3), np.eye(6) / 1e-2)
    lands = poses[:, 9:12].mean(axis=0) + rng.uniform(-15, 15, size=(3, 3))
    ri = np.array([0, 0, 0, 10, 11, 40, 40, 41, 88]); rl = np.array([0, 1, 0, 1, 1, 0, 0, 1, 0]); tau = rng.uniform(-0.03, cfg.dt + 0.03, size=len(ri))
    z = np.array([np.linalg.norm(lands[l] - synth._retract(POSE3, poses[i], vels[i] * t)[9:12]) for i, l, t in zip(ri, rl, tau)]) + rng.normal(size=len(ri)) * 0.1
    g.add_interp_range(ri, rl, z, np.full(len(ri), 0.1), np.full(len(ri), cfg.dt), tau)
    for l in range(3):
        g.add_prior_landmark(l, lands[l] + rng.normal(size=3) * 0.3, np.eye(3))       # landmark 2: prior only
    g.add_prior_pose(0, poses[0], np.eye(6) / 1e-3); g.add_prior_vel(0, vels[0], np.eye(6) / 1e-2)
    g.add_prior_vel(n - 1, vels[n - 1], np.eye(6) / 1e-1)
    for i in (4, 5, 46):  # states right behind a missing prior need their velocity tied down by something
        g.add_prior_vel(i, vels[i], np.eye(6))
    g.set_values(np.stack([synth._retract(POSE3, poses[i], rng.normal(size=6) * 0.03) for i in range(n)]), np.zeros((n, 6)), lands + rng.normal(size=lands.shape) * 0.3)
    if hasattr(g, "finalize"):
        g.finalize()
    return g


def test_ragged_graph_matches_oracle():
    g = _ragged_graph(lambda grp, n, l: gb.Graph(grp, n, l)); o = _ragged_graph(lambda grp, n, l: po.Graph(grp, n, l))
    assert abs(g.linearize() - o.error()) <= 1e-9 * o.error()
    Hg, gg = g.normal_equations_dense(); Ho, go = o.normal_equations_dense()
    np.testing.assert_allclose(Hg, Ho, atol=2e-6 * np.abs(Ho).max()); np.testing.assert_allclose(gg, go, atol=2e-6 * np.abs(go).max())
    for use_lm in (False, True):
        g = _ragged_graph(lambda grp, n, l: gb.Graph(grp, n, l)); o = _ragged_graph(lambda grp, n, l: po.Graph(grp, n, l))
        sg = g.optimize(use_lm=use_lm); so = o.optimize(use_lm=use_lm)
        assert sg.status == 0 and so.status == 0 and sg.iterations == so.iterations
        Pg, Vg, Lg = g.get_values(); Po, Vo, Lo = o.get_values()
        assert np.abs(Pg - Po).max() <= 1e-6 and np.abs(Vg - Vo).max() <= 1e-6 and np.abs(Lg - Lo).max() <= 1e-6


def test_indeterminate_system_is_reported():
    """a GP chain with no prior at all has a free gauge: Gauss-Newton must fail loudly (GTSAM throws
    IndeterminantLinearSystemException), LM must cope through its damping - in the engine as in the oracle"""
    from gpslam_b200 import capi
    n = 40
    cfg = small_cfg("C2", n); cfg.prior_every = 0
    poses, vels = synth.ground_truth(cfg)

    def build(make):
        g = make(POSE3, n, 0)
        g.add_qc_model(np.eye(6) * 0.01); g.add_gp_prior(np.arange(n - 1), np.full(n - 1, cfg.dt)); g.set_values(poses, vels * 1.01, None)
        if hasattr(g, "finalize"):
            g.finalize()
        return g
    o = build(lambda grp, nn, l: po.Graph(grp, nn, l))
    assert o.optimize(use_lm=False).status != 0
    g = build(lambda grp, nn, l: gb.Graph(grp, nn, l))
    with pytest.raises(capi.GpbError, match="indeterminate"):
        g.optimize(use_lm=False)
    g = build(lambda grp, nn, l: gb.Graph(grp, nn, l)); o = build(lambda grp, nn, l: po.Graph(grp, nn, l))
    sg = g.optimize(use_lm=True); so = o.optimize(use_lm=True)
    assert sg.status == 0 and so.status == 0 and sg.error_final < 1e-6 and so.error_final < 1e-6


@pytest.mark.parametrize("name,n_landmarks", [("C3", 17), ("C3", 18), ("C3", 38), ("C1", 60)])
def test_wide_landmark_borders(name, n_landmarks):
    """17 3-D landmarks is the widest border of the 64-column SE(3) production kernels; up to 38 3-D / 60 2-D landmarks run the
    128-column generic sweep (the reference has no landmark limit) - GN and LM against the oracle"""
    cfg = small_cfg(name, 150, n_landmarks=n_landmarks, prior_every=30, range_per_state=1.5)
    for use_lm in (False, True):
        g, o, _ = both(cfg)
        sg = g.optimize(use_lm=use_lm); so = o.optimize(use_lm=use_lm)
        assert sg.status == 0 and sg.iterations == so.iterations
        Pg, Vg, Lg = g.get_values(); Po, Vo, Lo = o.get_values()
        assert np.abs(Pg - Po).max() <= 1e-6 and np.abs(Vg - Vo).max() <= 1e-6 and np.abs(Lg - Lo).max() <= 1e-6


def test_landmark_border_limit():
    """beyond 128 - 2D - 1 border columns the graph is refused at finalize, not mis-solved"""
    from gpslam_b200 import capi
    for name, n_landmarks in (("C3", 39), ("C1", 61)):
        cfg = small_cfg(name, 150, n_landmarks=n_landmarks, prior_every=30, range_per_state=1.5)
        with pytest.raises(capi.GpbError):
            synth.build(cfg, lambda grp, n, l: gb.Graph(grp, n, l))


# ----------------------------------------------------------------------------- interpolatePose queries
@pytest.mark.parametrize("group", [POSE3, POSE2, ROT3, LINEAR, 4])
def test_interpolate_pose_queries(group):
    """gpb_interpolate_poses (GaussianProcessInterpolator*::interpolatePose, incl. Pose3VW = group 4) against the oracle: poses
    1e-11, Jacobians 1e-6 for SE(3) (the oracle keeps the reference's numerical differentiation) and 1e-10 otherwise"""
    from gpslam_b200 import capi
    from tests.test_hostmath_vs_oracle import rand_state
    rng = np.random.default_rng(1200 + group)
    base = POSE3 if group == 4 else group
    D = 6 if base == POSE3 else 3
    n = 48
    X1, V1, X2, V2, DT, TAU = [], [], [], [], [], []
    for k in range(n):
        p1, v1 = rand_state(rng, base)
        if k % 2:
            p2, v2 = rand_state(rng, base)
        else:
            step = rng.normal(size=D) * (1e-9 if k % 6 == 0 else 0.05)
            p2 = po.retract(base, p1, step) if base != POSE2 else po.pose2_compose(p1, po.pose2_expmap(step))
            v2 = v1 + rng.normal(size=D) * 0.1
        dt = float(rng.uniform(0.05, 0.5))
        X1.append(p1); V1.append(v1); X2.append(p2); V2.append(v2); DT.append(dt); TAU.append(float(rng.uniform(-0.5, 1.5) * dt))
    poses, H = capi.interpolate_poses(group, np.stack(X1), np.stack(V1), np.stack(X2), np.stack(V2), np.array(DT), np.array(TAU), want_H=True)
    only = capi.interpolate_poses(group, np.stack(X1), np.stack(V1), np.stack(X2), np.stack(V2), np.array(DT), np.array(TAU))
    assert np.array_equal(poses, only)
    for k in range(n):
        ref, Href = po.interpolate(group, np.eye(D), DT[k], TAU[k], X1[k], V1[k], X2[k], V2[k], want_H=True)
        np.testing.assert_allclose(poses[k], ref, atol=1e-11 * max(1.0, np.abs(ref).max()))
        for v in range(4):
            np.testing.assert_allclose(H[k, v], Href[v], atol=(1e-6 if D == 6 else 1e-10) * max(1.0, np.abs(Href[v]).max()))
    with pytest.raises(capi.GpbError):
        capi.interpolate_poses(group, X1[0], V1[0], X2[0], V2[0], -0.1, 0.0)


@pytest.mark.parametrize("case", ["pose3", "pose2", "rot3", "pose3vw"])
def test_graph_interpolate_dense_output(case):
    """gpb_graph_interpolate: the optimised trajectory queried between its states equals interpolatePose on the optimised values;
    tau = 0 and tau = delta_t reproduce the support states"""
    from gpslam_b200 import capi
    g, o = make_pair(case)
    g.optimize(use_lm=True)
    P, V, _ = g.get_values()
    rng = np.random.default_rng(5)
    iv = rng.integers(0, g.N - 1, size=40); dt = synth.config(CASES[case]["name"]).dt
    tau = rng.uniform(0, dt, size=40)
    got = g.interpolate(iv, tau)
    ref = capi.interpolate_poses(g.group, P[iv], V[iv], P[iv + 1], V[iv + 1], dt, tau)
    assert np.array_equal(got, ref)
    for k in range(0, 40, 7):
        r = po.interpolate(g.group, np.eye(g.D), dt, tau[k], P[iv[k]], V[iv[k]], P[iv[k] + 1], V[iv[k] + 1])
        np.testing.assert_allclose(got[k], r, atol=1e-11 * max(1.0, np.abs(r).max()))
    def same(a, b, tol):
        d = a - b
        if case == "pose2":   # headings compare modulo 2 pi (Logmap wraps the relative heading)
            d[:, 2] = np.arctan2(np.sin(d[:, 2]), np.cos(d[:, 2]))
        assert np.abs(d).max() <= tol * max(1.0, np.abs(P).max())
    same(g.interpolate(iv, np.zeros(40)), P[iv], 1e-12)
    same(g.interpolate(iv, np.full(40, dt)), P[iv + 1], 1e-9)
    with pytest.raises(capi.GpbError):
        g.interpolate([g.N - 1], [0.0])

