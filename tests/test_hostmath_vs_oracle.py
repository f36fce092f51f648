"""CPU check of the product's __host__ __device__ factor arithmetic (gpslam_b200/csrc/*.cuh, compiled by g++ through
tests/hostmath) against the oracle on random inputs: whitened GP-prior [A|b] for every group, interpolated range /
attitude rows.  Tolerances: residuals 1e-11; Jacobians 1e-6 where the oracle carries the reference's 1e-6-step numerical
differentiation (SE(3) blocks, gp/Pose3utils.cpp:167-179), 1e-10 elsewhere."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

from oracle import pyoracle as po

HERE = os.path.dirname(os.path.abspath(__file__))


@pytest.fixture(scope="module")
def hm():
    so = os.path.join(HERE, "hostmath", "libhostmath.so")
    subprocess.run(["g++", "-O2", "-std=c++17", "-fPIC", "-shared", "-x", "c++", "-o", so, os.path.join(HERE, "hostmath", "hostmath.cpp")], check=True)
    return C.CDLL(so)


def dp(a):
    return a.ctypes.data_as(C.POINTER(C.c_double))


def rand_state(rng, group, scale_rot=1.0, scale_t=3.0):
    if group == po.POSE3:
        T = po.pose3_expmap(np.concatenate([rng.normal(size=3) * scale_rot, rng.normal(size=3) * scale_t]))
        return T, rng.normal(size=6)
    if group == po.ROT3:
        return po.rot3_expmap(rng.normal(size=3) * scale_rot), rng.normal(size=3)
    if group == po.POSE2:
        return np.array([*(rng.normal(size=2) * scale_t), rng.normal() * scale_rot]), rng.normal(size=3)
    return rng.normal(size=3) * scale_t, rng.normal(size=3)


def rand_spd(rng, n):
    A = rng.normal(size=(n, n))
    return A @ A.T + n * np.eye(n)


@pytest.mark.parametrize("group", [po.POSE3, po.POSE2, po.ROT3, po.LINEAR])
@pytest.mark.parametrize("small", [False, True])
def test_gp_prior_whitened(hm, group, small):
    rng = np.random.default_rng(100 + group)
    D = 6 if group == po.POSE3 else 3
    m, ncol = 2 * D, 4 * D + 1
    for trial in range(40):
        dt = float(rng.uniform(0.05, 1.5))
        Qc = rand_spd(rng, D) if trial % 2 else np.eye(D) * rng.uniform(0.01, 2)
        p1, v1 = rand_state(rng, group)
        if small:  # trajectory-like: small relative motion (also theta -> 0 branches)
            step = rng.normal(size=D) * (1e-9 if trial % 5 == 0 else 0.05)
            p2 = po.retract(group, p1, step) if group != po.POSE2 else po.pose2_compose(p1, po.pose2_expmap(step))
            v2 = v1 + rng.normal(size=D) * 0.1
        else:
            p2, v2 = rand_state(rng, group)
        g = po.Graph(group, 2, 0)
        g.set_values(np.stack([p1, p2]), np.stack([v1, v2]))
        g.add_qc_model(Qc)
        g.add_gp_prior(0, dt)
        A, b = g.linearize_factor(0)
        Ao = np.concatenate(A + [b.reshape(-1, 1)], axis=1)
        Rq = np.linalg.cholesky(np.linalg.inv(Qc)).T  # upper, Rq^T Rq = Qc^-1
        out = np.zeros(m * ncol)
        s1 = np.concatenate([p1, v1]); s2 = np.concatenate([p2, v2])
        hm.hm_gp_prior(C.c_int(group), dp(s1), dp(s2), C.c_double(dt), dp(np.ascontiguousarray(Rq.T).ravel()), dp(out))
        Ag = out.reshape(ncol, m).T
        scale = max(1.0, np.abs(Ao).max())
        np.testing.assert_allclose(Ag[:, -1], Ao[:, -1], atol=1e-11 * scale)                      # rhs = -R e
        np.testing.assert_allclose(Ag[:, :-1], Ao[:, :-1], atol=(1e-6 if group == po.POSE3 else 1e-10) * scale)
        if group == po.POSE3:
            # the SE(3) kernel's instantiations agree with each other to rounding: struct form (101) always, diagonal-Rq form
            # (100) when Qc is diagonal
            for code in (101,) + ((100,) if trial % 2 == 0 else ()):
                out2 = np.zeros(m * ncol)
                hm.hm_gp_prior(C.c_int(code), dp(s1), dp(s2), C.c_double(dt), dp(np.ascontiguousarray(Rq.T).ravel()), dp(out2))
                np.testing.assert_allclose(out2, out, atol=1e-12 * scale, rtol=1e-12)


@pytest.mark.parametrize("group", [po.POSE3, po.POSE2, po.LINEAR])
def test_interp_range_rows(hm, group):
    rng = np.random.default_rng(7 + group)
    D = 6 if group == po.POSE3 else 3
    DL = 3 if group == po.POSE3 else 2
    for trial in range(40):
        dt = float(rng.uniform(0.05, 0.5)); tau = float(rng.uniform(-0.5, 1.5) * dt)
        p1, v1 = rand_state(rng, group)
        if trial % 2:
            p2, v2 = rand_state(rng, group)
        else:
            step = rng.normal(size=D) * 0.05
            p2 = po.retract(group, p1, step) if group != po.POSE2 else po.pose2_compose(p1, po.pose2_expmap(step))
            v2 = v1 + rng.normal(size=D) * 0.1
        land = rng.normal(size=DL) * 10
        sensor = None
        if trial % 3 == 0 and group != po.LINEAR:
            sensor = rand_state(rng, group, 0.5, 0.5)[0]
        g = po.Graph(group, 2, 1)
        g.set_values(np.stack([p1, p2]), np.stack([v1, v2]), land.reshape(1, -1))
        g.add_qc_model(np.eye(D))
        z = float(rng.uniform(1, 20))
        g.add_interp_range(0, 0, z, 0.1, dt, tau, body_P_sensor=sensor)
        e, H = g.eval_factor(0, True)
        prm = np.zeros(56); prm[0] = dt; prm[1] = tau; prm[2] = z
        if sensor is not None:
            prm[4:4 + len(sensor)] = sensor; prm[16] = 1.0
        out = np.zeros(4 * D + DL + 1)
        hm.hm_interp_range(C.c_int(group), dp(np.concatenate([p1, v1])), dp(np.concatenate([p2, v2])), dp(land), dp(prm), dp(out))
        Ho = np.concatenate([h.ravel() for h in H])
        scale = max(1.0, np.abs(Ho).max())
        assert abs(out[-1] - e[0]) < 1e-10 * max(1.0, abs(e[0]))
        np.testing.assert_allclose(out[:-1], Ho, atol=(1e-6 if group == po.POSE3 else 1e-10) * scale)


def test_interp_attitude_rows(hm):
    rng = np.random.default_rng(11)
    for trial in range(40):
        dt = float(rng.uniform(0.005, 0.5)); tau = float(rng.uniform(0, 1) * dt)
        p1, v1 = rand_state(rng, po.ROT3)
        p2 = po.retract(po.ROT3, p1, rng.normal(size=3) * (0.05 if trial % 2 else 1.0)); v2 = v1 + rng.normal(size=3) * 0.1
        nz = rng.normal(size=3); nz /= np.linalg.norm(nz)
        br = rng.normal(size=3); br /= np.linalg.norm(br)
        if trial % 4 == 0:
            nz, br = np.array([0, 0, 1.0]), np.array([0, 0, 1.0])
        g = po.Graph(po.ROT3, 2, 0)
        g.set_values(np.stack([p1, p2]), np.stack([v1, v2]))
        g.add_qc_model(np.eye(3))
        g.add_interp_attitude(0, dt, tau, nz, 0.1, bRef=br)
        e, H = g.eval_factor(0, True)
        prm = np.zeros(56); prm[0] = dt; prm[1] = tau; prm[4:7] = nz; prm[7:10] = br
        out = np.zeros(26)
        hm.hm_interp_attitude(dp(np.concatenate([p1, v1])), dp(np.concatenate([p2, v2])), dp(prm), dp(out))
        for r in range(2):
            assert abs(out[13 * r + 12] - e[r]) < 1e-11
            Ho = np.concatenate([h[r] for h in H])
            np.testing.assert_allclose(out[13 * r:13 * r + 12], Ho, atol=1e-10 * max(1.0, np.abs(Ho).max()))


@pytest.mark.parametrize("small", [False, True])
def test_gp_prior_vw_whitened(hm, small):
    """GaussianProcessPriorPose3VW: velocities [v_world | w_world]; same 12x25 [A|b] shape as the body-velocity prior"""
    rng = np.random.default_rng(321)
    for trial in range(40):
        dt = float(rng.uniform(0.05, 1.5))
        Qc = rand_spd(rng, 6) if trial % 2 else np.eye(6) * rng.uniform(0.01, 2)
        p1, v1 = rand_state(rng, po.POSE3)
        if small:
            p2 = po.retract(po.POSE3, p1, rng.normal(size=6) * (1e-9 if trial % 5 == 0 else 0.05)); v2 = v1 + rng.normal(size=6) * 0.1
        else:
            p2, v2 = rand_state(rng, po.POSE3)
        g = po.Graph(po.POSE3VW, 2, 0)
        g.set_values(np.stack([p1, p2]), np.stack([v1, v2]))
        g.add_qc_model(Qc)
        g.add_gp_prior(0, dt)
        A, b = g.linearize_factor(0)
        Ao = np.concatenate(A + [b.reshape(-1, 1)], axis=1)
        Rq = np.linalg.cholesky(np.linalg.inv(Qc)).T
        out = np.zeros(12 * 25)
        hm.hm_gp_prior_vw(dp(np.concatenate([p1, v1])), dp(np.concatenate([p2, v2])), C.c_double(dt), dp(np.ascontiguousarray(Rq.T).ravel()), dp(out))
        Ag = out.reshape(25, 12).T
        scale = max(1.0, np.abs(Ao).max())
        np.testing.assert_allclose(Ag[:, -1], Ao[:, -1], atol=1e-11 * scale)
        np.testing.assert_allclose(Ag[:, :-1], Ao[:, :-1], atol=1e-6 * scale)


@pytest.mark.parametrize("vw", [0, 1])
def test_interp_gps_rows(hm, vw):
    """GPInterpolatedGPSFactorPose3 / ...Pose3VW rows (the oracle carries the reference's numerical differentiation: 1e-6)"""
    rng = np.random.default_rng(55 + vw)
    for trial in range(40):
        dt = float(rng.uniform(0.05, 0.5)); tau = float(rng.uniform(-0.5, 1.5) * dt)
        p1, v1 = rand_state(rng, po.POSE3)
        if trial % 2:
            p2, v2 = rand_state(rng, po.POSE3)
        else:
            p2 = po.retract(po.POSE3, p1, rng.normal(size=6) * 0.05); v2 = v1 + rng.normal(size=6) * 0.1
        sensor = rand_state(rng, po.POSE3, 0.5, 0.5)[0] if trial % 3 == 0 else None
        meas = rng.normal(size=3) * 5
        g = po.Graph(po.POSE3VW if vw else po.POSE3, 2, 0)
        g.set_values(np.stack([p1, p2]), np.stack([v1, v2]))
        g.add_qc_model(np.eye(6))
        g.add_interp_gps(0, meas, np.eye(3), dt, tau, body_P_sensor=sensor)
        e, H = g.eval_factor(0, True)
        prm = np.zeros(56); prm[0] = dt; prm[1] = tau; prm[40:43] = meas
        if sensor is not None:
            prm[4:16] = sensor; prm[16] = 1.0
        out = np.zeros(75)
        hm.hm_interp_gps(C.c_int(vw), dp(np.concatenate([p1, v1])), dp(np.concatenate([p2, v2])), dp(prm), dp(out))
        for r in range(3):
            assert abs(out[25 * r + 24] - e[r]) < 1e-10 * max(1.0, abs(e[r]))
            Ho = np.concatenate([h[r] for h in H])
            np.testing.assert_allclose(out[25 * r:25 * r + 24], Ho, atol=1e-6 * max(1.0, np.abs(Ho).max()))


def test_interp_projection_rows(hm):
    rng = np.random.default_rng(77)
    K = np.array([50.0, 45.0, 0.3, 40.0, 30.0])
    for trial in range(40):
        dt = float(rng.uniform(0.05, 0.5)); tau = float(rng.uniform(-0.5, 1.5) * dt)
        p1, v1 = rand_state(rng, po.POSE3, 0.3)
        p2 = po.retract(po.POSE3, p1, rng.normal(size=6) * 0.05); v2 = v1 + rng.normal(size=6) * 0.1
        sensor = rand_state(rng, po.POSE3, 0.2, 0.5)[0] if trial % 3 == 0 else None
        R, t = po.pose3_Rt(p1)
        land = t + R @ np.array([rng.normal(), rng.normal(), rng.uniform(-3, 12)])  # some behind the camera: cheirality branch
        meas = rng.normal(size=2) * 20
        g = po.Graph(po.POSE3, 2, 1)
        g.set_values(np.stack([p1, p2]), np.stack([v1, v2]), land.reshape(1, 3))
        g.add_qc_model(np.eye(6))
        g.add_interp_projection(0, 0, meas, np.eye(2), dt, tau, K, body_P_sensor=sensor)
        e, H = g.eval_factor(0, True)
        prm = np.zeros(56); prm[0] = dt; prm[1] = tau; prm[40:42] = meas; prm[43:48] = K
        if sensor is not None:
            prm[4:16] = sensor; prm[16] = 1.0
        out = np.zeros(56)
        hm.hm_interp_projection(dp(np.concatenate([p1, v1])), dp(np.concatenate([p2, v2])), dp(land), dp(prm), dp(out))
        for r in range(2):
            assert abs(out[28 * r + 27] - e[r]) < 1e-9 * max(1.0, abs(e[r]))
            Ho = np.concatenate([h[r] for h in H])
            np.testing.assert_allclose(out[28 * r:28 * r + 27], Ho, atol=1e-6 * max(1.0, np.abs(Ho).max()))


@pytest.mark.parametrize("group", [po.POSE3, po.POSE2, po.ROT3, po.LINEAR, po.POSE3VW])
def test_interpolate_pose_query(hm, group):
    """interpolatePose as a query (gpb_interpolate_poses / GaussianProcessInterpolator*::interpolatePose): pose and Hint1..4"""
    rng = np.random.default_rng(900 + group)
    base = po.POSE3 if group == po.POSE3VW else group
    D = 6 if base == po.POSE3 else 3
    PS = {po.POSE3: 12, po.POSE2: 3, po.ROT3: 9, po.LINEAR: 3}[base]
    for trial in range(40):
        dt = float(rng.uniform(0.05, 0.5)); tau = float(rng.uniform(-0.5, 1.5) * dt)
        p1, v1 = rand_state(rng, base)
        if trial % 2:
            p2, v2 = rand_state(rng, base)
        else:
            step = rng.normal(size=D) * (1e-9 if trial % 6 == 0 else 0.05)
            p2 = po.retract(base, p1, step) if base != po.POSE2 else po.pose2_compose(p1, po.pose2_expmap(step))
            v2 = v1 + rng.normal(size=D) * 0.1
        ref, Href = po.interpolate(group, np.eye(D), dt, tau, p1, v1, p2, v2, want_H=True)
        out = np.zeros(PS); H = np.zeros(4 * D * D)
        hm.hm_interp_pose(C.c_int(group), dp(np.concatenate([p1, v1])), dp(np.concatenate([p2, v2])), C.c_double(dt), C.c_double(tau), dp(out), dp(H))
        np.testing.assert_allclose(out, ref, atol=1e-11 * max(1.0, np.abs(ref).max()))
        for k in range(4):
            Hk = H[k * D * D:(k + 1) * D * D].reshape(D, D).T
            np.testing.assert_allclose(Hk, Href[k], atol=(1e-6 if D == 6 else 1e-10) * max(1.0, np.abs(Href[k]).max()))
        out2 = np.zeros(PS)
        hm.hm_interp_pose(C.c_int(group), dp(np.concatenate([p1, v1])), dp(np.concatenate([p2, v2])), C.c_double(dt), C.c_double(tau), dp(out2), None)
        assert np.array_equal(out, out2)     # pose-only path


def test_se3_prior_and_interpolator_by_rotation_regime(hm):
    """The SE(3) prior's [A|b] and interpolatePose across the regimes of the relative rotation angle theta between the two states:
    1e-7 (series branches), 1e-4 and 1e-6 (around the reference's theta = 1e-5 switch of rightJacobianPose3Q, gp/Pose3utils.cpp:98,
    where its 1e-6-step numerical differentiation is itself only good to ~1e-6), 0.3 (typical), 2.5-3.1 (large).  Residuals and
    poses agree with the oracle to rounding everywhere; Jacobians to 1e-8 where the reference's numerical derivative is clean and to
    2e-6 around the switch.  (theta within 1e-3 of pi is excluded: Logmap is ill-conditioned there in the reference as well, and
    adjacent states of a trajectory are never half a turn apart.)"""
    rng = np.random.default_rng(2024)
    tolJ = {0: 1e-8, 1: 2e-6, 2: 1e-8, 3: 1e-7, 4: 2e-6}
    for trial in range(500):
        mode = trial % 5
        dt = float(rng.uniform(0.01, 1.0))
        p1, v1 = rand_state(rng, po.POSE3, scale_rot=2.0)
        if mode == 0:
            step = rng.normal(size=6) * 1e-7
        elif mode == 1:
            step = rng.normal(size=6) * 1e-4
        elif mode == 2:
            step = rng.normal(size=6) * 0.3
        elif mode == 3:
            ax = rng.normal(size=3); ax /= np.linalg.norm(ax)
            step = np.concatenate([ax * rng.uniform(2.5, 3.1), rng.normal(size=3) * 5])
        else:
            step = np.concatenate([rng.normal(size=3) * 1e-6, rng.normal(size=3)])
        p2 = po.retract(po.POSE3, p1, step); v2 = v1 + rng.normal(size=6) * (0.1 if mode < 3 else 2.0)
        g = po.Graph(po.POSE3, 2, 0); g.set_values(np.stack([p1, p2]), np.stack([v1, v2])); g.add_qc_model(np.eye(6)); g.add_gp_prior(0, dt)
        A, b = g.linearize_factor(0)
        Ao = np.concatenate(A + [b.reshape(-1, 1)], axis=1)
        out = np.zeros(12 * 25); s1 = np.concatenate([p1, v1]); s2 = np.concatenate([p2, v2])
        hm.hm_gp_prior(C.c_int(0), dp(s1), dp(s2), C.c_double(dt), dp(np.eye(6).ravel()), dp(out))
        Ag = out.reshape(25, 12).T; sc = max(1.0, np.abs(Ao).max())
        assert np.abs(Ag[:, -1] - Ao[:, -1]).max() <= 1e-11 * sc, (trial, mode)
        assert np.abs(Ag[:, :-1] - Ao[:, :-1]).max() <= tolJ[mode] * sc, (trial, mode)
        tau = float(rng.uniform(-0.5, 1.5) * dt)
        ref, Href = po.interpolate(po.POSE3, np.eye(6), dt, tau, p1, v1, p2, v2, want_H=True)
        o2 = np.zeros(12); H = np.zeros(144)
        hm.hm_interp_pose(C.c_int(0), dp(s1), dp(s2), C.c_double(dt), C.c_double(tau), dp(o2), dp(H))
        assert np.abs(o2 - ref).max() <= 1e-11 * max(1.0, np.abs(ref).max()), (trial, mode)
        for k in range(4):
            Hk = H[k * 36:(k + 1) * 36].reshape(6, 6).T
            assert np.abs(Hk - Href[k]).max() <= tolJ[mode] * max(1.0, np.abs(Href[k]).max()), (trial, mode, k)
