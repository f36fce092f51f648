"""Deterministic synthetic trajectory graphs for BASELINE.json's configs.

`build(cfg, make_graph)` constructs the same graph through any object exposing the construction calls shared by
gpslam_b200.Graph (CUDA engine) and oracle.pyoracle.Graph (CPU oracle), so both sides see identical inputs.
Pure numpy; no arithmetic of the hot path lives here (ground truth uses its own small Exp maps).
"""
import numpy as np

POSE3, POSE2, ROT3, LINEAR = 0, 1, 2, 3
POSE3VW = 4  # graph group of SE(3) states with [v_world | w_world] velocities (Config.vw); cfg.group stays POSE3


def _skew(w):
    return np.array([[0, -w[2], w[1]], [w[2], 0, -w[0]], [-w[1], w[0], 0]])


def _so3_exp(w):
    th = np.linalg.norm(w)
    W = _skew(w)
    if th < 1e-12:
        return np.eye(3) + W
    return np.eye(3) + np.sin(th) / th * W + (1 - np.cos(th)) / th ** 2 * (W @ W)


def _se3_exp(xi):
    w, v = xi[:3], xi[3:]
    R = _so3_exp(w)
    th2 = w @ w
    if th2 < 1e-20:
        return R, v.copy()
    wxv = np.array([w[1] * v[2] - w[2] * v[1], w[2] * v[0] - w[0] * v[2], w[0] * v[1] - w[1] * v[0]])  # np.cross, without its per-call overhead
    return R, (wxv - R @ wxv + w * (w @ v)) / th2


def _se2_exp(xi):
    w = xi[2]
    if abs(w) < 1e-12:
        return xi.copy()
    c, s = np.cos(w), np.sin(w)
    ox, oy = -xi[1], xi[0]
    return np.array([(ox - (c * ox - s * oy)) / w, (oy - (s * ox + c * oy)) / w, w])


def _wire3(R, t):
    return np.concatenate([R.T.ravel(), t])


class Config:
    def __init__(self, name, group, n_states, n_landmarks=0, range_per_state=0.0, dt=0.1, seed=0, qc_sigma=0.1, prior_every=100,
                 attitude_every=0, odometry=False, init_noise=0.05, zero_rot_fraction=0.01, n_closures=0, closure_min_gap=0, closure_ends=False, gps_every=0, proj_per_state=0.0, vw=False, qc_dense=False):
        self.__dict__.update(locals())
        del self.__dict__["self"]


def config(name):
    """BASELINE.json configs (sizes may be overridden by the caller through attributes)."""
    base = 20260925
    if name == "C1":
        return Config("C1", POSE2, 1000, 4, 0.444, dt=0.1, seed=base + 1, qc_sigma=0.1, prior_every=0, odometry=True)
    if name == "C2":
        return Config("C2", POSE3, 10000, 0, 0.0, dt=0.1, seed=base + 2, qc_sigma=0.1)
    if name == "C3":
        return Config("C3", POSE3, 100000, 16, 0.5, dt=0.1, seed=base + 3, qc_sigma=0.1)
    if name == "C4":
        return Config("C4", ROT3, 1000000, 0, 0.0, dt=0.005, seed=base + 4, qc_sigma=100.0, prior_every=0, attitude_every=4)
    if name == "VW":
        # SE(3) "VW" family: GaussianProcessPriorPose3VW chain + GPInterpolatedGPSFactorPose3VW fixes; the
        # reference has no range / projection factor for these states, so no landmarks
        return Config("VW", POSE3, 10000, 0, 0.0, dt=0.1, seed=base + 6, qc_sigma=0.1, gps_every=5, vw=True)
    if name == "C5":
        # K = 128 loop closures (BetweenFactor<Pose3>, sigma 0.05) between random state pairs >= 10 000 states apart
        return Config("C5", POSE3, 1000000, 16, 0.5, dt=0.1, seed=base + 5, qc_sigma=0.1, n_closures=128, closure_min_gap=10000)
    raise KeyError(name)


def ground_truth(cfg):
    """constant body twist with a slow sinusoidal modulation: T_{i+1} = T_i Exp(w_i dt), v_i = w_i"""
    rng = np.random.default_rng(cfg.seed)
    N, dt = cfg.n_states, cfg.dt
    t = np.arange(N) * dt
    if cfg.group == POSE3:
        w0 = np.array([0.05, -0.1, 0.27, 1.0, 0.1, -0.05])
        tw = w0[None, :] * (1 + 0.3 * np.sin(0.05 * t)[:, None]) + 0.02 * np.stack([np.sin(0.11 * t + k) for k in range(6)], axis=1)
        nz = max(1, int(cfg.zero_rot_fraction * N))
        z0 = N // 3
        tw[z0:z0 + nz, :3] = 0.0  # slice with exactly zero angular rate: hits the reference's theta -> 0 branches
        poses = np.zeros((N, 12)); R = np.eye(3); p = np.zeros(3)
        for i in range(N):
            poses[i] = _wire3(R, p)
            dR, dp = _se3_exp(tw[i] * dt)
            p = R @ dp + p; R = R @ dR
            if i % 1000 == 0:
                u, _, vt = np.linalg.svd(R); R = u @ vt
        return poses, tw
    if cfg.group == ROT3:
        w0 = np.array([0.3, -0.2, 0.5])
        tw = w0[None, :] * (1 + 0.3 * np.sin(0.5 * t)[:, None])
        poses = np.zeros((N, 9)); R = np.eye(3)
        for i in range(N):
            poses[i] = R.T.ravel()
            R = R @ _so3_exp(tw[i] * dt)
            if i % 1000 == 0:
                u, _, vt = np.linalg.svd(R); R = u @ vt
        return poses, tw
    # planar groups
    w0 = np.array([1.0, 0.0, 0.12])
    tw = w0[None, :] * (1 + 0.3 * np.sin(0.05 * t)[:, None])
    poses = np.zeros((N, 3)); x = np.zeros(3)
    for i in range(N):
        poses[i] = x
        d = _se2_exp(tw[i] * dt)
        c, s = np.cos(x[2]), np.sin(x[2])
        x = np.array([x[0] + c * d[0] - s * d[1], x[1] + s * d[0] + c * d[1], x[2] + d[2]])
    if cfg.group == LINEAR:  # "linear Pose2" states: velocities are world-frame derivatives of (x, y, theta)
        v = np.zeros_like(tw)
        v[:, 0] = np.cos(poses[:, 2]) * tw[:, 0] - np.sin(poses[:, 2]) * tw[:, 1]
        v[:, 1] = np.sin(poses[:, 2]) * tw[:, 0] + np.cos(poses[:, 2]) * tw[:, 1]
        v[:, 2] = tw[:, 2]
        return poses, v
    return poses, tw


def _retract(group, pose, d):
    if group == POSE3:
        R = pose[:9].reshape(3, 3).T; t = pose[9:]
        dR, dp = _se3_exp(d)
        return _wire3(R @ dR, R @ dp + t)
    if group == ROT3:
        return (pose.reshape(3, 3).T @ _so3_exp(d)).T.ravel()
    if group == POSE2:
        c, s = np.cos(pose[2]), np.sin(pose[2])
        return np.array([pose[0] + c * d[0] - s * d[1], pose[1] + s * d[0] + c * d[1], pose[2] + d[2]])
    return pose + d


def build(cfg, make_graph, finalize=True):
    """Returns (graph, truth dict).  make_graph(group, n_states, n_landmarks) -> Graph-like object."""
    rng = np.random.default_rng(cfg.seed + 1000)
    N, dt, group = cfg.n_states, cfg.dt, cfg.group
    D = 6 if group == POSE3 else 3
    DL = {POSE3: 3, POSE2: 2, ROT3: 0, LINEAR: 2}[group]
    poses, vels = ground_truth(cfg)
    L = cfg.n_landmarks if DL else 0
    vw = bool(getattr(cfg, "vw", False)) and group == POSE3
    g = make_graph(POSE3VW if vw else group, N, L)
    # wire velocities: the body twist (w, v), or for VW states [v_world | w_world] = [R v | R w] (gp/Pose3utils.cpp:27-45)
    wire_vels = vels
    if vw:
        Rs = poses[:, :9].reshape(N, 3, 3).transpose(0, 2, 1)
        wire_vels = np.concatenate([np.einsum("nij,nj->ni", Rs, vels[:, 3:]), np.einsum("nij,nj->ni", Rs, vels[:, :3])], axis=1)
    Qc = np.eye(D) * cfg.qc_sigma ** 2
    if getattr(cfg, "qc_dense", False):  # a correlated Qc: the dense-Rq whitening path of the linearise kernels
        B = np.fromfunction(lambda i, j: 0.3 / (1.0 + np.abs(i - j)), (D, D)) + np.eye(D) * 0.7
        Qc = cfg.qc_sigma ** 2 * (B @ B.T)
    g.add_qc_model(Qc)
    g.add_gp_prior(np.arange(N - 1), np.full(N - 1, dt))
    iso = lambda n, s: np.eye(n) / s
    lands = np.zeros((L, max(DL, 1)))
    if L and vw:
        raise ValueError("synth: Pose3 VW graphs carry no landmarks (no range / projection factor in the reference)")
    if L:
        if group == POSE3:
            ctr = poses[:, 9:12].mean(axis=0); span = np.abs(poses[:, 9:12] - ctr).max() + 10.0
            lands = ctr + rng.uniform(-span, span, size=(L, 3))
        else:
            ctr = poses[:, :2].mean(axis=0); span = np.abs(poses[:, :2] - ctr).max() + 10.0
            lands = ctr + rng.uniform(-span, span, size=(L, 2))
        nr = int(round(cfg.range_per_state * N))
        ri = np.sort(rng.integers(0, N - 1, size=nr)); rl = rng.integers(0, L, size=nr); tau = rng.uniform(0, dt, size=nr)
        sig = 0.1 if group == POSE3 else 0.5
        z = np.zeros(nr)
        for k in range(nr):
            i = ri[k]
            Ti = _retract(group if group != LINEAR else POSE2, poses[i], (vels[i] if group != LINEAR else _body(poses[i], vels[i])) * tau[k])
            if group == POSE3:
                z[k] = np.linalg.norm(lands[rl[k]] - Ti[9:12])
            else:
                z[k] = np.linalg.norm(lands[rl[k]] - Ti[:2])
        z += rng.normal(size=nr) * sig
        g.add_interp_range(ri, rl, z, np.full(nr, sig), np.full(nr, dt), tau)
        for l in range(L):
            g.add_prior_landmark(l, lands[l] + rng.normal(size=DL) * 0.5, iso(DL, 1.0))
    # SE(3) only: interpolated GPS fixes (every gps_every-th interval, sensor offset, correlated noise model) and interpolated
    # pinhole projections of the landmarks (Cal3_S2 with skew); SURVEY.md §8f rank 2.  Own generator: the rest of the graph does
    # not change when they are switched on.
    if group == POSE3 and (cfg.gps_every or cfg.proj_per_state):
        mrng = np.random.default_rng(cfg.seed + 3000)
        bTs = _wire3(_so3_exp(np.array([0.1, -0.2, 0.3])), np.array([0.3, 0.6, -0.7]))
        def sensor_pose(i, tau_):
            Ti = _retract(POSE3, poses[i], vels[i] * tau_)
            R = Ti[:9].reshape(3, 3).T; t = Ti[9:]
            Rs = bTs[:9].reshape(3, 3).T; ts = bTs[9:]
            return R @ Rs, R @ ts + t
        if cfg.gps_every:
            gi = np.arange(0, N - 1, cfg.gps_every); gtau = mrng.uniform(0, dt, size=len(gi))
            gmeas = np.stack([sensor_pose(i, t_)[1] for i, t_ in zip(gi, gtau)]) + mrng.normal(size=(len(gi), 3)) * 0.05
            g.add_interp_gps(gi, gmeas, np.array([[20.0, 2.0, -1.0], [0.0, 15.0, 3.0], [0.0, 0.0, 25.0]]), np.full(len(gi), dt), gtau, body_P_sensor=bTs)
        if cfg.proj_per_state and L:
            K = np.array([50.0, 45.0, 0.5, 40.0, 30.0])
            pi_, pl_, ptau, pmeas = [], [], [], []
            behind = None
            for i in mrng.integers(0, N - 1, size=int(round(cfg.proj_per_state * N)) * 4):
                tau_ = float(mrng.uniform(0, dt)); l = int(mrng.integers(0, L))
                Rc, tc = sensor_pose(int(i), tau_)
                q = Rc.T @ (lands[l] - tc)
                if q[2] < -5.0 and behind is None:
                    behind = (int(i), l, tau_)   # one factor whose landmark is behind the camera: the cheirality path
                if q[2] > 5.0 and len(pi_) < int(round(cfg.proj_per_state * N)):
                    u, v = q[0] / q[2], q[1] / q[2]
                    pi_.append(int(i)); pl_.append(l); ptau.append(tau_)
                    pmeas.append([K[0] * u + K[2] * v + K[3] + mrng.normal() * 0.3, K[1] * v + K[4] + mrng.normal() * 0.3])
            if behind is not None:
                pi_.append(behind[0]); pl_.append(behind[1]); ptau.append(behind[2]); pmeas.append([40.0, 30.0])
            if pi_:
                g.add_interp_projection(np.array(pi_), np.array(pl_), np.array(pmeas), np.array([[3.0, 0.4], [0.0, 2.5]]), np.full(len(pi_), dt), np.array(ptau), K, body_P_sensor=bTs)
    # gauge: pose + velocity prior on state 0, sparse pose priors along the chain
    s0 = 1e-3 if group != POSE2 else 1.0
    g.add_prior_pose(0, poses[0], iso(D, s0))
    g.add_prior_vel(0, wire_vels[0], iso(D, 1e-3 if group != POSE2 else 1.0))
    if cfg.prior_every:
        for i in range(cfg.prior_every, N, cfg.prior_every):
            g.add_prior_pose(i, _retract(group, poses[i], rng.normal(size=D) * 0.1), iso(D, 0.1))
    if cfg.attitude_every and group == ROT3:
        ai = np.arange(0, N - 1, cfg.attitude_every)
        g.add_interp_attitude(ai, np.full(len(ai), dt), np.full(len(ai), 0.5 * dt), _attitude_meas(poses, vels, ai, dt, rng), np.full(len(ai), 0.1))
        # gyro-like relative-rotation factors (the reference's AHRSFactor is GTSAM code)
    if cfg.odometry and group in (POSE2, POSE3, ROT3):
        for i in range(N - 1):
            meas = _between(group, poses[i], poses[i + 1], rng, 1e-3)
            g.add_between(i, i + 1, meas, iso(D, 1e-3) if group != POSE2 else np.diag([1e3, 1e3, 1e3 / np.pi]))
    # loop closures: BetweenFactor between distant states (GTSAM type; C5).  Uses its own generator so that the rest of the graph
    # (and the initial values below) do not depend on the number of closures.
    if cfg.n_closures and group in (POSE2, POSE3, ROT3):
        crng = np.random.default_rng(cfg.seed + 2000)
        gap = min(cfg.closure_min_gap if cfg.closure_min_gap else max(2, N // 10), max(2, N - 2))
        for k in range(cfg.n_closures):
            i = int(crng.integers(0, N - gap)); j = int(crng.integers(i + gap, N))
            if cfg.closure_ends and k == 0:
                i, j = 0, N - 1  # first and last state of the chain as endpoints
            if cfg.closure_ends and k == 1:
                i, j = 1, N - 2  # endpoints adjacent to other endpoints (segments without interior)
            if crng.random() < 0.25:
                i, j = j, i  # also exercise the (later -> earlier) argument order
            meas = _between(group, poses[i], poses[j], crng, 0.05)
            g.add_between(i, j, meas, iso(D, 0.05))
    # initial values: truth (+) noise on the pose tangent, zero velocities (matlab/PlazaPose2.m:201-202)
    init = np.stack([_retract(group, poses[i], rng.normal(size=D) * cfg.init_noise) for i in range(N)])
    g.set_values(init, np.zeros((N, D)), lands + (rng.normal(size=lands.shape) * 0.5 if L else 0))
    if finalize and hasattr(g, "finalize"):
        g.finalize()
    return g, dict(poses=poses, vels=wire_vels, lands=lands)


class Recorder:
    """Records the construction calls of build() once so that the same graph can be replayed into several graph objects (the CUDA
    engine and the CPU oracle of a parity test) without running the generator - a Python loop over every state - twice."""

    def __init__(self, group, n_states, n_landmarks):
        self.group, self.n_states, self.n_landmarks, self.calls = group, n_states, n_landmarks, []

    def __getattr__(self, name):
        if name.startswith("_") or name == "finalize":
            raise AttributeError(name)
        def rec(*a, **kw):
            self.calls.append((name, a, kw))
            return sum(1 for c in self.calls if c[0] == name) - 1   # add_qc_model returns the model id
        return rec

    def replay(self, make_graph, finalize=True):
        g = make_graph(self.group, self.n_states, self.n_landmarks)
        for name, a, kw in self.calls:
            getattr(g, name)(*a, **kw)
        if finalize and hasattr(g, "finalize"):
            g.finalize()
        return g


def record(cfg):
    """(Recorder, truth) of build(cfg, ...): replay it into as many graph objects as needed"""
    box = []
    def mk(grp, n, l):
        box.append(Recorder(grp, n, l)); return box[0]
    _, truth = build(cfg, mk, finalize=False)
    return box[0], truth


def _body(pose, v):
    c, s = np.cos(pose[2]), np.sin(pose[2])
    return np.array([c * v[0] + s * v[1], -s * v[0] + c * v[1], v[2]])


def _between(group, a, b, rng, noise):
    if group == POSE3:
        Ra = a[:9].reshape(3, 3).T; Rb = b[:9].reshape(3, 3).T
        R = Ra.T @ Rb; t = Ra.T @ (b[9:] - a[9:])
        dR, dp = _se3_exp(rng.normal(size=6) * noise)
        return _wire3(R @ dR, R @ dp + t)
    if group == ROT3:
        return ((a.reshape(3, 3).T).T @ b.reshape(3, 3).T @ _so3_exp(rng.normal(size=3) * noise)).T.ravel()
    c, s = np.cos(a[2]), np.sin(a[2])
    dx, dy = b[0] - a[0], b[1] - a[1]
    return np.array([c * dx + s * dy, -s * dx + c * dy, b[2] - a[2]]) + rng.normal(size=3) * noise


def _attitude_meas(poses, vels, ai, dt, rng):
    """accelerometer-style direction measurements nZ = R(t) bRef + noise, bRef = e_z"""
    out = np.zeros((len(ai), 3))
    for k, i in enumerate(ai):
        R = poses[i].reshape(3, 3).T @ _so3_exp(vels[i] * 0.5 * dt)
        n = R @ np.array([0, 0, 1.0]) + rng.normal(size=3) * 0.02
        out[k] = n / np.linalg.norm(n)
    return out
