"""On-disk formats and example-graph builders (gpslam_b200/datasets.py): Plaza .mat and IMU / MOCAP text
tables.  CPU: round trips through the file formats, the builders on the oracle against known ground truth, and - when the
reference checkout is present in this container - the reference's own datasets end to end.  GPU: the engine against the oracle
on a Plaza-format file written by the test (nothing under /root/reference is read on the GPU box)."""
import os

import numpy as np
import pytest

from gpslam_b200 import datasets as ds
from gpslam_b200 import synth

REF_DATA = "/root/reference/matlab/data"


def synthetic_plaza(path, n=300, L=4, seed=3):
    """a Plaza-shaped log from the C1 ground truth: odometry (forward, heading) per step, biased ranges to L beacons at
    asynchronous times, beacon ids that are not 0..L-1 (as in the real logs)"""
    rng = np.random.default_rng(seed)
    cfg = synth.config("C1"); cfg.n_states = n
    poses, vels = synth.ground_truth(cfg)
    t = 3152.0 + np.cumsum(np.concatenate([[0.0], rng.uniform(0.08, 0.12, size=n - 1)]))
    # re-integrate the truth on the irregular time stamps so that odometry and truth agree
    GT = np.zeros((n, 4)); GT[:, 0] = t
    x = np.zeros(3); DR = np.zeros((n - 1, 3))
    for i in range(n):
        GT[i, 1:] = x
        if i + 1 < n:
            dt = t[i + 1] - t[i]
            fwd, dth = vels[i, 0] * dt, vels[i, 2] * dt
            DR[i] = [t[i + 1], fwd + rng.normal() * 1e-3, dth + rng.normal() * 1e-3]
            x = np.array([x[0] + np.cos(x[2]) * fwd, x[1] + np.sin(x[2]) * fwd, x[2] + dth])
    ids = [1, 6, 0, 5][:L]
    ctr = GT[:, 1:3].mean(axis=0)
    TL = np.array([[ids[k], *(ctr + rng.uniform(-30, 30, size=2))] for k in range(L)])
    m = int(0.45 * n)
    tm = np.sort(rng.uniform(t[0] + 0.01, t[-1], size=m))
    TD = np.zeros((m, 4))
    for k in range(m):
        i = np.searchsorted(t, tm[k]) - 1
        a = (tm[k] - t[i]) / (t[i + 1] - t[i])
        p = (1 - a) * GT[i, 1:3] + a * GT[i + 1, 1:3]
        l = int(rng.integers(0, L))
        true = np.linalg.norm(TL[l, 1:3] - p)
        TD[k] = [tm[k], 2, ids[l], (true - 0.02) / 0.93 + rng.normal() * 0.3]   # biased the way the real beacons are
    TD[5, 3] += 25.0   # one gross outlier for the mask
    ds.save_plaza(path, GT, DR, TL, TD, init_heading_offset=0.0)
    return GT, TL, TD


def test_plaza_roundtrip_and_fit(tmp_path):
    p = str(tmp_path / "plaza_synth.mat")
    GT, TL, TD = synthetic_plaza(p)
    d = ds.load_plaza(p)
    assert np.array_equal(d["GT"], GT) and np.array_equal(d["TL"], TL) and np.array_equal(d["TD"], TD) and d["init_heading_offset"] == 0.0
    trans, mask = ds.range_measure_fit(d["GT"], d["TL"], d["TD"])
    assert mask[5] and mask.sum() <= 6                 # the planted outlier is found
    assert abs(trans[0] - 0.93) < 0.02 and abs(trans[1] - 0.02) < 0.5


@pytest.mark.parametrize("use_linear", [False, True])
def test_plaza_graph_on_oracle(tmp_path, use_linear):
    from oracle import pyoracle as po
    p = str(tmp_path / "plaza_synth.mat")
    synthetic_plaza(p)
    d = ds.load_plaza(p)
    g, info = ds.build_plaza(d, lambda grp, n, l: po.Graph(grp, n, l), use_linear=use_linear)
    assert info["n_poses"] == 300 and info["n_outliers"] >= 1 and info["n_ranges_used"] + info["n_outliers"] == len(d["TD"])
    assert g.num_factors() == 4 + 1 + 299 + 299 + info["n_ranges_used"]   # landmark priors, first-pose prior, odometry, GP priors, ranges
    e0 = g.error()
    st = g.optimize(use_lm=True)
    assert st.status == 0 and st.error_final < e0
    P, _, _ = g.get_values()
    pos, rot = ds.plaza_errors(d, P)
    assert pos < 0.25 and rot < 0.05


@pytest.mark.skipif(not os.path.exists(os.path.join(REF_DATA, "Plaza2.mat")), reason="reference datasets not present (GPU box)")
def test_reference_plaza2_end_to_end():
    """matlab/PlazaPose2.m on its own dataset through the oracle: 4091 poses, 1816 ranges, LM to convergence"""
    from oracle import pyoracle as po
    d = ds.load_plaza(os.path.join(REF_DATA, "Plaza2.mat"))
    g, info = ds.build_plaza(d, lambda grp, n, l: po.Graph(grp, n, l))
    g.set_threads(po.hardware_threads())
    assert info["n_poses"] == 4091 and info["n_ranges_used"] + info["n_outliers"] == 1816
    st = g.optimize(use_lm=True)
    assert st.status == 0
    P, _, _ = g.get_values()
    pos, rot = ds.plaza_errors(d, P)
    assert pos < 0.3 and rot < 0.05, (pos, rot)      # 0.15 m / 0.017 rad when this was written


def synthetic_imu(tmp_path, n=2000, seed=4):
    rng = np.random.default_rng(seed)
    t = np.arange(n) * 0.005
    w = np.stack([0.3 * np.sin(0.7 * t), 0.2 * np.cos(0.5 * t), 0.4 * np.ones(n)], axis=1)
    R = np.eye(3); Rs = []
    for k in range(n):
        if k:
            R = R @ ds._so3_exp(w[k] * (t[k] - t[k - 1]))
        Rs.append(R.copy())
    acc = np.stack([Rk.T @ np.array([0, 0, 9.81]) for Rk in Rs]) + rng.normal(size=(n, 3)) * 0.05
    imu = np.concatenate([np.arange(n)[:, None] + 100, t[:, None], w + rng.normal(size=(n, 3)) * 1e-4, acc], axis=1)
    p = str(tmp_path / "imu.txt"); np.savetxt(p, imu, fmt="%.9f")
    # quaternion x y z w of every 3rd pose
    q = []
    for Rk in Rs[::3]:
        qw = 0.5 * np.sqrt(max(0.0, 1 + np.trace(Rk)))
        q.append([(Rk[2, 1] - Rk[1, 2]) / (4 * qw), (Rk[0, 2] - Rk[2, 0]) / (4 * qw), (Rk[1, 0] - Rk[0, 1]) / (4 * qw), qw])
    moc = np.concatenate([np.arange(len(q))[:, None], t[::3, None], np.zeros((len(q), 3)), np.array(q)], axis=1)
    pm = str(tmp_path / "mocap.txt"); np.savetxt(pm, moc, fmt="%.9f")
    return p, pm, np.array(Rs)


def test_ahrs_tables_and_graph_on_oracle(tmp_path):
    from oracle import pyoracle as po
    p, pm, Rs = synthetic_imu(tmp_path)
    imu = ds.load_imu_txt(p); moc, Rm = ds.load_mocap_txt(pm)
    assert imu.shape == (2000, 8) and moc.shape[1] == 9
    np.testing.assert_allclose(Rm[5].reshape(3, 3).T, Rs[15], atol=1e-7)          # quaternion -> wire rotation
    # gyro_dt / acc_dt a hair under the sample period: with exact 5 ms stamps, t[m] - t[m-1] >= 0.005 fails half the time in floating point
    g, info = ds.build_ahrs(imu, Rm[0], lambda grp, n, l: po.Graph(grp, n, l), gyro_dt=0.0049, acc_dt=0.0199, max_time=9.0)
    assert info["n_states"] > 1500 and info["n_acc"] > 300
    st = g.optimize(use_lm=True)
    assert st.status == 0
    P, _, _ = g.get_values()
    for k in range(0, info["n_states"], 97):
        Ra = P[k].reshape(3, 3).T; Rb = Rs[info["state_meas_idx"][k]]
        ang = np.degrees(np.arccos(np.clip((np.trace(Ra.T @ Rb) - 1) / 2, -1, 1)))
        assert ang < 0.5, (k, ang)


@pytest.mark.skipif(not os.path.exists(os.path.join(REF_DATA, "RAW_IMU_DATA_matlab.txt")), reason="reference datasets not present (GPU box)")
def test_reference_ahrs_dataset_loads_and_solves():
    from oracle import pyoracle as po
    imu = ds.load_imu_txt(os.path.join(REF_DATA, "RAW_IMU_DATA_matlab.txt"))
    moc, Rm = ds.load_mocap_txt(os.path.join(REF_DATA, "MOCAP_POSE_DATA_matlab.txt"))
    assert imu.shape == (9834, 8) and moc.shape == (7128, 9)
    g, info = ds.build_ahrs(imu, Rm[0], lambda grp, n, l: po.Graph(grp, n, l), max_time=10.0)
    g.set_threads(po.hardware_threads())
    e0 = g.error(); st = g.optimize(use_lm=True)
    assert st.status == 0 and st.error_final <= e0 and info["n_states"] > 1000


@pytest.mark.gpu
@pytest.mark.parametrize("use_linear", [False, True])
def test_plaza_engine_matches_oracle(tmp_path, use_linear):
    import gpslam_b200 as gb
    from oracle import pyoracle as po
    p = str(tmp_path / "plaza_synth.mat")
    synthetic_plaza(p)
    d = ds.load_plaza(p)
    g, _ = ds.build_plaza(d, lambda grp, n, l: gb.Graph(grp, n, l), use_linear=use_linear)
    o, _ = ds.build_plaza(d, lambda grp, n, l: po.Graph(grp, n, l), use_linear=use_linear)
    assert abs(g.linearize() - o.error()) <= 1e-9 * o.error()
    sg = g.optimize(use_lm=True); so = o.optimize(use_lm=True)
    assert sg.status == 0 and sg.iterations == so.iterations
    Pg, Vg, Lg = g.get_values(); Po, Vo, Lo = o.get_values()
    assert np.abs(Pg - Po).max() <= 1e-6 and np.abs(Vg - Vo).max() <= 1e-6 and np.abs(Lg - Lo).max() <= 1e-6
    pos, rot = ds.plaza_errors(d, Pg)
    assert pos < 0.25


@pytest.mark.gpu
def test_ahrs_engine_matches_oracle(tmp_path):
    import gpslam_b200 as gb
    from oracle import pyoracle as po
    p, pm, _ = synthetic_imu(tmp_path, n=1200)
    imu = ds.load_imu_txt(p); _, Rm = ds.load_mocap_txt(pm)
    g, _ = ds.build_ahrs(imu, Rm[0], lambda grp, n, l: gb.Graph(grp, n, l), gyro_dt=0.0049, acc_dt=0.0199, max_time=5.5)
    o, _ = ds.build_ahrs(imu, Rm[0], lambda grp, n, l: po.Graph(grp, n, l), gyro_dt=0.0049, acc_dt=0.0199, max_time=5.5)
    assert abs(g.linearize() - o.error()) <= 1e-9 * max(1.0, o.error())
    sg = g.optimize(use_lm=True); so = o.optimize(use_lm=True)
    assert sg.status == 0 and sg.iterations == so.iterations
    Pg, Vg, _ = g.get_values(); Po, Vo, _ = o.get_values()
    assert np.abs(Pg - Po).max() <= 1e-6 and np.abs(Vg - Vo).max() <= 1e-6
