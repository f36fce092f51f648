"""On-disk formats of the reference's examples and the graph construction of its two MATLAB scripts,
so the datasets the reference ships (matlab/data/*.mat, *.txt) run end to end through the engine without MATLAB.

  load_plaza(path)                 matlab/PlazaPose2.m:13-24   Plaza1/2 .mat (GT, DR, DRp, TL, TD, init_heading_offset)
  range_measure_fit(GT, TL, TD)    matlab/range_measure_fit.m:1-93  range bias fit + outlier mask
  build_plaza(data, make_graph)    matlab/PlazaPose2.m:26-204  Pose2 (or 2DLinear) GP-prior + odometry + interpolated-range graph
  plaza_errors(data, poses)        matlab/PlazaPose2.m:238-262 mean position / heading error against ground truth
  load_imu_txt / load_mocap_txt    matlab/GPAHRSexample.m:42-63 RAW_IMU_DATA / MOCAP_POSE_DATA text tables
  build_ahrs(imu, att, make_graph) matlab/GPAHRSexample.m:66-214 Rot3 GP-prior + interpolated-attitude graph (see its docstring for
                                   what stands in for GTSAM's AHRSFactor)

Host-side data plumbing only (numpy / scipy.io): every factor goes through the construction calls shared by
gpslam_b200.Graph (CUDA engine) and oracle.pyoracle.Graph (CPU oracle); no arithmetic of the hot path lives here.
"""
import numpy as np

POSE3, POSE2, ROT3, LINEAR = 0, 1, 2, 3


# ----------------------------------------------------------------------------------------------------------------- Plaza
def load_plaza(path):
    """Plaza .mat (MATLAB v5): GT [n x 4] (t, x, y, heading), DR [n-1 x 3] (t, forward odometry, heading odometry), TL [L x 3]
    (landmark id, x, y), TD [m x 4] (t, sender, landmark id, range), init_heading_offset (matlab/PlazaPose2.m:13-19)."""
    import scipy.io as sio
    d = sio.loadmat(path)
    out = {k: np.asarray(d[k], dtype=np.float64) for k in ("GT", "DR", "TL", "TD") if k in d}
    missing = [k for k in ("GT", "DR", "TL", "TD") if k not in out]
    if missing:
        raise ValueError("load_plaza: %s lacks %s" % (path, ", ".join(missing)))
    if "DRp" in d:
        out["DRp"] = np.asarray(d["DRp"], dtype=np.float64)
    out["init_heading_offset"] = float(np.asarray(d["init_heading_offset"]).ravel()[0]) if "init_heading_offset" in d else 0.0
    return out


def save_plaza(path, GT, DR, TL, TD, init_heading_offset=0.0):
    """writes the same .mat layout (tests and synthetic stand-ins for the datasets)"""
    import scipy.io as sio
    sio.savemat(path, {"GT": np.asarray(GT, float), "DR": np.asarray(DR, float), "TL": np.asarray(TL, float), "TD": np.asarray(TD, float),
                       "init_heading_offset": np.array([[float(init_heading_offset)]])})


def range_measure_fit(GT, TL, TD, outlier_limit=2.0):
    """matlab/range_measure_fit.m: associates every range with the nearer of the two ground-truth poses around it, fits
    true = a * measured + b by least squares, masks |fit - true| > 2 m as outliers and refits on the inliers.
    Returns (range_trans (a, b), outlier_mask bool [m])."""
    T = GT[:, 0]
    nr_pose, nr_range = len(GT), len(TD)
    lid = {int(l): k for k, l in enumerate(TL[:, 0])}
    pose_of = np.zeros(nr_range, dtype=np.int64)
    i, v = 0, 0
    while i < nr_pose and v < nr_range:  # :30-48 (0-based here)
        t = TD[v, 0]
        if T[i] >= t:
            pose_of[v] = i - 1 if (i >= 1 and abs(T[i - 1] - t) <= abs(T[i] - t)) else i
            v += 1
        else:
            i += 1
    pose_of[v:] = nr_pose - 1  # measurements after the last pose (the MATLAB loop leaves them at index 0; harmless there, fixed here)
    lands = np.array([[TL[lid[int(l)], 1], TL[lid[int(l)], 2]] for l in TD[:, 2]])
    true_range = np.linalg.norm(GT[pose_of, 1:3] - lands, axis=1)
    meas = TD[:, 3]
    A = np.stack([meas, np.ones(nr_range)], axis=1)
    x = np.linalg.lstsq(A, true_range, rcond=None)[0]
    mask = np.abs(x[0] * meas + x[1] - true_range) > outlier_limit
    x = np.linalg.lstsq(A[~mask], true_range[~mask], rcond=None)[0]
    return x, mask


def _pose2_compose(a, b):
    c, s = np.cos(a[2]), np.sin(a[2])
    return np.array([a[0] + c * b[0] - s * b[1], a[1] + s * b[0] + c * b[1], a[2] + b[2]])


def build_plaza(data, make_graph, use_linear=False, add_odometry=True, init_ground_truth=False, add_first_pose_prior=True,
                add_first_vel_zero_prior=False, add_landmark_prior=True, max_poses=None, finalize=True):
    """matlab/PlazaPose2.m:26-204 with the script's settings as defaults (Qc sigma 0.1, first-pose prior (1, 1, pi), odometry
    (1, 1, pi) * 1e-3, range sigma 0.5, landmark prior sigma 1).  States are the ground-truth time stamps; a range measured at
    time t joins the interval (i-1, i) whose right end is the first pose with time >= t, tau = t - t_{i-1}.
    Returns (graph, info) with info = dict(n_poses, n_ranges_used, n_outliers, land_ids, range_trans)."""
    GT, DR, TL, TD = data["GT"], data["DR"], data["TL"], data["TD"]
    off = data.get("init_heading_offset", 0.0)
    n = len(GT) if max_poses is None else min(len(GT), int(max_poses))
    L = len(TL)
    range_trans, outlier = range_measure_fit(GT, TL, TD)
    group = LINEAR if use_linear else POSE2
    g = make_graph(group, n, L)
    g.add_qc_model(np.eye(3) * 0.1 ** 2)
    iso = lambda sig: np.diag(1.0 / np.asarray(sig, dtype=np.float64))
    lid = {int(l): k for k, l in enumerate(TL[:, 0])}
    for k in range(L):
        if add_landmark_prior:
            g.add_prior_landmark(k, TL[k, 1:3], iso([1.0, 1.0]))
    first = np.array([GT[0, 1], GT[0, 2], GT[0, 3] + off])
    if add_first_pose_prior:
        g.add_prior_pose(0, first, iso([1.0, 1.0, np.pi]))
    if add_first_vel_zero_prior:
        g.add_prior_vel(0, np.zeros(3), iso([1.0, 1.0, np.pi]))
    odom_R = iso(np.array([1.0, 1.0, np.pi]) * 1e-3)
    init = np.zeros((n, 3)); init[0] = first
    last_pose, last_vec = first.copy(), first.copy()
    ri, rl, rz, rdt, rtau = [], [], [], [], []
    nxt = 0
    nr_range = len(TD)
    dts = np.zeros(n - 1)
    for p in range(1, n):
        delta_t = GT[p, 0] - GT[p - 1, 0]
        dts[p - 1] = delta_t
        odom = np.array([DR[p - 1, 1], 0.0, DR[p - 1, 2]])
        new_pose = _pose2_compose(last_pose, odom)
        new_vec = last_vec + np.array([new_pose[0] - last_pose[0], new_pose[1] - last_pose[1], odom[2]])
        if add_odometry:
            if use_linear:
                g.add_odometry_2d(p - 1, p, odom, odom_R)          # OdometryFactor2DLinear (:121-122)
            else:
                g.add_between(p - 1, p, odom, odom_R)              # BetweenFactorPose2 (:124-125)
        last_pose, last_vec = new_pose, new_vec
        while nxt < nr_range and TD[nxt, 0] <= GT[p, 0]:           # :146-177
            if not outlier[nxt]:
                ri.append(p - 1); rl.append(lid[int(TD[nxt, 2])]); rz.append(range_trans[0] * TD[nxt, 3] + range_trans[1])
                rdt.append(delta_t); rtau.append(TD[nxt, 0] - GT[p - 1, 0])
            nxt += 1
        init[p] = (np.array([GT[p, 1], GT[p, 2], GT[p, 3] + off]) if init_ground_truth else (last_vec if use_linear else last_pose))
    g.add_gp_prior(np.arange(n - 1), dts)                          # GaussianProcessPrior{Pose2,Linear3} (:130-139)
    if ri:
        g.add_interp_range(np.array(ri), np.array(rl), np.array(rz), np.full(len(ri), 0.5), np.array(rdt), np.array(rtau))
    g.set_values(init, np.zeros((n, 3)), TL[:, 1:3].copy())
    if finalize and hasattr(g, "finalize"):
        g.finalize()
    return g, dict(n_poses=n, n_ranges_used=len(ri), n_outliers=int(outlier[:nxt].sum()), land_ids=[int(l) for l in TL[:, 0]], range_trans=range_trans,
                   group=group)


def plaza_errors(data, poses):
    """matlab/PlazaPose2.m:238-262: mean Euclidean position error and mean |heading error| (wrapped) against ground truth"""
    GT = data["GT"][:len(poses)]
    off = data.get("init_heading_offset", 0.0)
    pos = np.linalg.norm(poses[:, :2] - GT[:, 1:3], axis=1)
    rot = poses[:, 2] - (GT[:, 3] + off)
    rot = np.arctan2(np.sin(rot), np.cos(rot))
    return float(pos.mean()), float(np.abs(rot).mean())


# ----------------------------------------------------------------------------------------------------------------- AHRS
def load_imu_txt(path):
    """RAW_IMU_DATA_matlab.txt: seq, time (s), gyro x/y/z (rad/s), acc x/y/z (m/s^2)  (matlab/GPAHRSexample.m:42-43)"""
    a = np.loadtxt(path)
    if a.ndim != 2 or a.shape[1] < 8:
        raise ValueError("load_imu_txt: expected >= 8 columns (seq, t, gyro xyz, acc xyz)")
    return a


def load_mocap_txt(path):
    """MOCAP_POSE_DATA_matlab.txt: seq, time (s), position x/y/z, orientation quaternion x/y/z/w (matlab/GPAHRSexample.m:44-52);
    returns (table, rotations [n x 9] in the engine's column-major wire layout)"""
    a = np.loadtxt(path)
    if a.ndim != 2 or a.shape[1] < 9:
        raise ValueError("load_mocap_txt: expected >= 9 columns (seq, t, pos xyz, quat xyzw)")
    x, y, z, w = a[:, 5], a[:, 6], a[:, 7], a[:, 8]
    nrm = np.sqrt(x * x + y * y + z * z + w * w)
    x, y, z, w = x / nrm, y / nrm, z / nrm, w / nrm
    R = np.empty((len(a), 3, 3))
    R[:, 0, 0] = 1 - 2 * (y * y + z * z); R[:, 0, 1] = 2 * (x * y - z * w); R[:, 0, 2] = 2 * (x * z + y * w)
    R[:, 1, 0] = 2 * (x * y + z * w); R[:, 1, 1] = 1 - 2 * (x * x + z * z); R[:, 1, 2] = 2 * (y * z - x * w)
    R[:, 2, 0] = 2 * (x * z - y * w); R[:, 2, 1] = 2 * (y * z + x * w); R[:, 2, 2] = 1 - 2 * (x * x + y * y)
    return a, R.transpose(0, 2, 1).reshape(len(a), 9).copy()


def _so3_exp(w):
    th = np.linalg.norm(w)
    W = np.array([[0, -w[2], w[1]], [w[2], 0, -w[0]], [-w[1], w[0], 0]])
    if th < 1e-12:
        return np.eye(3) + W
    return np.eye(3) + np.sin(th) / th * W + (1 - np.cos(th)) / th ** 2 * (W @ W)


def build_ahrs(imu, first_rotation, make_graph, gyro_dt=0.005, acc_dt=0.02, max_time=50.0, use_gyro=True, use_acc=True, finalize=True):
    """matlab/GPAHRSexample.m:66-214: states at the gyro rate (a new Rot3 state whenever gyro_dt has elapsed), accelerometer
    samples at acc_dt as GPInterpolatedAttitudeFactorRot3(nZ = (0,0,1), bRef = measured acceleration direction) on the interval
    they fall in, GaussianProcessPriorRot3 (Qc sigma 100) between consecutive states, PriorFactor<Rot3> (sigma 0.1) on the first.
    The script's gyro factors are GTSAM's AHRSFactor over pre-integrated measurements with a bias state per pose - GTSAM code,
    outside gpslam.  Here the pre-integrated rotation (product of Exp(gyro * dt), zero bias) enters as a
    BetweenFactor<Rot3> with the script's gyro sigma (1e-4 per axis).  An accelerometer sample that coincides with a state's
    time stamp (the script's non-interpolated Rot3AttitudeFactor, :163-168) is added as the interpolated factor with tau = dt.
    Initial rotations: the integrated gyro (what the script's gyro-only pre-solve converges to).  Returns (graph, info)."""
    t = imu[:, 1]
    states, rel, acc_list = [0], [], []   # measurement index of every state; relative rotations; (interval, tau, dt, acc)
    cached = []
    last_gyro_t = t[0]; last_acc_t = t[0] - acc_dt
    pim = np.eye(3)
    m = 0
    while m < len(imu) and t[m] < max_time:
        if m > 0:
            delta = t[m] - t[m - 1]
        if use_acc and t[m] - last_acc_t >= acc_dt:
            cached.append(m); last_acc_t = t[m]
        if m > 0 and (m == len(imu) - 1 or t[m] - last_gyro_t >= gyro_dt):
            pim = pim @ _so3_exp(imu[m, 2:5] * delta)
            dt = t[m] - last_gyro_t
            k = len(states) - 1
            rel.append((pim.copy(), dt))
            for a in cached:
                acc_list.append((k, t[a] - last_gyro_t, dt, imu[m, 5:8].copy()))  # the script passes the CURRENT sample's acceleration (:167, :176)
            cached = []
            pim = np.eye(3); last_gyro_t = t[m]; states.append(m)
        elif m > 0:
            pim = pim @ _so3_exp(imu[m, 2:5] * delta)
        m += 1
    n = len(states)
    if n < 2:
        raise ValueError("build_ahrs: fewer than two states (check gyro_dt / max_time)")
    g = make_graph(ROT3, n, 0)
    g.add_qc_model(np.eye(3) * 100.0 ** 2)
    g.add_gp_prior(np.arange(n - 1), np.array([d for _, d in rel]))
    R0 = np.asarray(first_rotation, dtype=np.float64).reshape(9)
    g.add_prior_pose(0, R0, np.eye(3) / 0.1)
    init = np.zeros((n, 9)); R = R0.reshape(3, 3).T.copy(); init[0] = R.T.ravel()
    for k, (dR, _) in enumerate(rel):
        if use_gyro:
            g.add_between(k, k + 1, dR.T.ravel(), np.eye(3) / 1e-4)
        R = R @ dR
        init[k + 1] = R.T.ravel()
    if acc_list:
        ai = np.array([a[0] for a in acc_list]); tau = np.array([a[1] for a in acc_list]); dts = np.array([a[2] for a in acc_list])
        b = np.stack([a[3] / np.linalg.norm(a[3]) for a in acc_list])
        g.add_interp_attitude(ai, dts, tau, np.tile([0.0, 0.0, 1.0], (len(ai), 1)), np.full(len(ai), 0.1), bRef=b)
    g.set_values(init, np.zeros((n, 3)), None)
    if finalize and hasattr(g, "finalize"):
        g.finalize()
    return g, dict(n_states=n, n_acc=len(acc_list), state_meas_idx=np.array(states), state_times=t[np.array(states)])
