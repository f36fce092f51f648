"""ORACLE — TEST INFRASTRUCTURE ONLY.

ctypes binding of oracle/libgpo.so (the CPU restatement of the reference's hot path).
Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
import this module, and only as the checker / CPU baseline — never the product path.
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None

# POSE3VW: Pose3 with [v_world; w_world] velocities (the reference's "VW" family); the C side keeps it as a POSE3 graph whose
# GP-prior / GPS factors are the VW kinds, exactly like the engine's GPB_POSE3VW
POSE3, POSE2, ROT3, LINEAR, POSE3VW = 0, 1, 2, 3, 4
POSE_STORAGE = {POSE3: 12, POSE2: 3, ROT3: 9, POSE3VW: 12}
TANGENT_DIM = {POSE3: 6, POSE2: 3, ROT3: 3, POSE3VW: 6}
LANDMARK_DIM = {POSE3: 3, POSE2: 2, ROT3: 0, LINEAR: 2, POSE3VW: 3}


def build():
    """Compile oracle/libgpo.so with the committed Makefile (no-op when up to date)."""
    subprocess.run(["make", "-s", "-C", _HERE], check=True)


def _dp(a):
    return a.ctypes.data_as(C.POINTER(C.c_double)) if a is not None else None


def _ip(a):
    return a.ctypes.data_as(C.POINTER(C.c_int)) if a is not None else None


def _f64(a):
    return np.ascontiguousarray(np.asarray(a, dtype=np.float64))


def _fcol(a):
    """matrix -> flat column-major float64"""
    return np.ascontiguousarray(np.asarray(a, dtype=np.float64).T).ravel()


class Params(C.Structure):
    _fields_ = [("max_iterations", C.c_int), ("rel_tol", C.c_double), ("abs_tol", C.c_double), ("err_tol", C.c_double),
                ("lambda_initial", C.c_double), ("lambda_factor", C.c_double), ("lambda_upper", C.c_double),
                ("lambda_lower", C.c_double), ("min_model_fidelity", C.c_double), ("use_lm", C.c_int)]


class Stats(C.Structure):
    _fields_ = [("iterations", C.c_int), ("error_initial", C.c_double), ("error_final", C.c_double), ("lambda_", C.c_double),
                ("lin_seconds", C.c_double), ("solve_seconds", C.c_double), ("total_seconds", C.c_double), ("status", C.c_int)]


def lib():
    global _LIB
    if _LIB is None:
        path = os.path.join(_HERE, "libgpo.so")
        if not os.path.exists(path):
            build()
        L = C.CDLL(path)
        L.gpo_graph_create.restype = C.c_void_p
        L.gpo_graph_create.argtypes = [C.c_int, C.c_int, C.c_int, C.c_int]
        L.gpo_error.restype = C.c_double
        for name in ("gpo_graph_destroy", "gpo_set_threads", "gpo_add_qc_model", "gpo_add_gp_prior", "gpo_add_interp_range",
                     "gpo_add_interp_attitude", "gpo_add_interp_gps", "gpo_add_interp_projection", "gpo_add_gp_prior_vw", "gpo_add_interp_gps_vw", "gpo_add_prior_pose", "gpo_add_prior_vel", "gpo_add_prior_landmark", "gpo_add_between",
                     "gpo_add_range_2d", "gpo_add_range_bearing_2d", "gpo_add_odometry_2d", "gpo_set_values", "gpo_get_values",
                     "gpo_num_factors", "gpo_check_step", "gpo_error", "gpo_linearize_factor", "gpo_eval_factor", "gpo_normal_equations_dense", "gpo_optimize"):
            getattr(L, name).argtypes = None
        _LIB = L
    return _LIB


def default_params(use_lm=True):
    p = Params()
    lib().gpo_default_params(C.byref(p), C.c_int(1 if use_lm else 0))
    return p


class Graph:
    """Oracle-side trajectory factor graph (same construction calls as gpslam_b200.Graph)."""

    def __init__(self, group, n_states, n_landmarks=0, dim=3):
        self.L = lib()
        self.group, self.N = group, n_states
        self.D = TANGENT_DIM.get(group, dim)
        self.PS = POSE_STORAGE.get(group, dim)
        self.DL = LANDMARK_DIM[group]
        self.NL = n_landmarks if self.DL else 0
        self.vw = group == POSE3VW
        self.h = C.c_void_p(self.L.gpo_graph_create(POSE3 if self.vw else group, dim, n_states, n_landmarks))

    def __del__(self):
        if getattr(self, "h", None):
            self.L.gpo_graph_destroy(self.h)
            self.h = None

    def set_threads(self, t):
        self.L.gpo_set_threads(self.h, C.c_int(t))

    def add_qc_model(self, Qc):
        return self.L.gpo_add_qc_model(self.h, _dp(_fcol(Qc)))

    def add_gp_prior(self, i, delta_t, qc=0):
        i = np.ascontiguousarray(np.atleast_1d(i), dtype=np.int32)
        dt = _f64(np.broadcast_to(np.atleast_1d(delta_t), i.shape))
        if self.vw:
            return self.add_gp_prior_vw(i, delta_t, qc)
        self.L.gpo_add_gp_prior(self.h, C.c_int(len(i)), _ip(i), _dp(dt), C.c_int(qc))

    def add_interp_range(self, i, l, z, sigma, delta_t, tau, qc=0, body_P_sensor=None):
        i = np.ascontiguousarray(np.atleast_1d(i), dtype=np.int32)
        l = np.ascontiguousarray(np.broadcast_to(np.atleast_1d(l), i.shape), dtype=np.int32)
        b = lambda a: _f64(np.broadcast_to(np.atleast_1d(a), i.shape))
        bps = _f64(body_P_sensor) if body_P_sensor is not None else None
        assert not self.vw, "no range factor for Pose3 VW states in the reference"
        self.L.gpo_add_interp_range(self.h, C.c_int(len(i)), _ip(i), _ip(l), _dp(b(z)), _dp(b(sigma)), _dp(b(delta_t)), _dp(b(tau)), C.c_int(qc), _dp(bps))

    def add_interp_gps(self, i, meas, sqrt_info, delta_t, tau, qc=0, body_P_sensor=None):
        """GPInterpolatedGPSFactorPose3: meas [n x 3] points, sqrt_info 3x3 upper-triangular R shared by the n factors"""
        if self.vw:
            return self.add_interp_gps_vw(i, meas, sqrt_info, delta_t, tau, qc, body_P_sensor)
        i = np.ascontiguousarray(np.atleast_1d(i), dtype=np.int32)
        b = lambda a: _f64(np.broadcast_to(np.atleast_1d(a), i.shape))
        m = _f64(np.broadcast_to(np.asarray(meas, dtype=np.float64).reshape(-1, 3), (len(i), 3)))
        bps = _f64(body_P_sensor) if body_P_sensor is not None else None
        rc = self.L.gpo_add_interp_gps(self.h, C.c_int(len(i)), _ip(i), _dp(m), _dp(_fcol(sqrt_info)), _dp(b(delta_t)), _dp(b(tau)), C.c_int(qc), _dp(bps))
        assert rc == 0

    def add_gp_prior_vw(self, i, delta_t, qc=0):
        """GaussianProcessPriorPose3VW: the graph's velocities are [v_world; w_world] (Pose3 graphs only)"""
        i = np.ascontiguousarray(np.atleast_1d(i), dtype=np.int32)
        dt = _f64(np.broadcast_to(np.atleast_1d(delta_t), i.shape))
        rc = self.L.gpo_add_gp_prior_vw(self.h, C.c_int(len(i)), _ip(i), _dp(dt), C.c_int(qc))
        assert rc == 0

    def add_interp_gps_vw(self, i, meas, sqrt_info, delta_t, tau, qc=0, body_P_sensor=None):
        """GPInterpolatedGPSFactorPose3VW"""
        i = np.ascontiguousarray(np.atleast_1d(i), dtype=np.int32)
        b = lambda a: _f64(np.broadcast_to(np.atleast_1d(a), i.shape))
        m = _f64(np.broadcast_to(np.asarray(meas, dtype=np.float64).reshape(-1, 3), (len(i), 3)))
        bps = _f64(body_P_sensor) if body_P_sensor is not None else None
        rc = self.L.gpo_add_interp_gps_vw(self.h, C.c_int(len(i)), _ip(i), _dp(m), _dp(_fcol(sqrt_info)), _dp(b(delta_t)), _dp(b(tau)), C.c_int(qc), _dp(bps))
        assert rc == 0

    def add_interp_projection(self, i, l, meas, sqrt_info, delta_t, tau, K, qc=0, body_P_sensor=None):
        """GPInterpolatedProjectionFactorPose3<Cal3_S2>: meas [n x 2] image points, K = (fx, fy, s, u0, v0)"""
        i = np.ascontiguousarray(np.atleast_1d(i), dtype=np.int32)
        l = np.ascontiguousarray(np.broadcast_to(np.atleast_1d(l), i.shape), dtype=np.int32)
        b = lambda a: _f64(np.broadcast_to(np.atleast_1d(a), i.shape))
        m = _f64(np.broadcast_to(np.asarray(meas, dtype=np.float64).reshape(-1, 2), (len(i), 2)))
        bps = _f64(body_P_sensor) if body_P_sensor is not None else None
        rc = self.L.gpo_add_interp_projection(self.h, C.c_int(len(i)), _ip(i), _ip(l), _dp(m), _dp(_fcol(sqrt_info)), _dp(b(delta_t)), _dp(b(tau)), C.c_int(qc),
                                              _dp(_f64(K)), _dp(bps))
        assert rc == 0

    def add_interp_attitude(self, i, delta_t, tau, nZ, sigma, bRef=(0, 0, 1), qc=0):
        i = np.ascontiguousarray(np.atleast_1d(i), dtype=np.int32)
        b = lambda a: _f64(np.broadcast_to(np.atleast_1d(a), i.shape))
        nz = _f64(np.broadcast_to(np.asarray(nZ, dtype=np.float64).reshape(-1, 3), (len(i), 3)))
        br = _f64(np.broadcast_to(np.asarray(bRef, dtype=np.float64).reshape(-1, 3), (len(i), 3)))
        self.L.gpo_add_interp_attitude(self.h, C.c_int(len(i)), _ip(i), _dp(b(delta_t)), _dp(b(tau)), C.c_int(qc), _dp(nz), _dp(br), _dp(b(sigma)))

    def add_prior_pose(self, i, value, sqrt_info):
        self.L.gpo_add_prior_pose(self.h, C.c_int(i), _dp(_f64(value)), _dp(_fcol(sqrt_info)))

    def add_prior_vel(self, i, value, sqrt_info):
        self.L.gpo_add_prior_vel(self.h, C.c_int(i), _dp(_f64(value)), _dp(_fcol(sqrt_info)))

    def add_prior_landmark(self, l, value, sqrt_info):
        self.L.gpo_add_prior_landmark(self.h, C.c_int(l), _dp(_f64(value)), _dp(_fcol(sqrt_info)))

    def add_between(self, i, j, meas, sqrt_info):
        self.L.gpo_add_between(self.h, C.c_int(i), C.c_int(j), _dp(_f64(meas)), _dp(_fcol(sqrt_info)))

    def add_range_2d(self, i, l, z, sigma):
        self.L.gpo_add_range_2d(self.h, C.c_int(i), C.c_int(l), C.c_double(z), C.c_double(sigma))

    def add_range_bearing_2d(self, i, l, rng, bearing, sqrt_info):
        self.L.gpo_add_range_bearing_2d(self.h, C.c_int(i), C.c_int(l), C.c_double(rng), C.c_double(bearing), _dp(_fcol(sqrt_info)))

    def add_odometry_2d(self, i, j, meas, sqrt_info):
        self.L.gpo_add_odometry_2d(self.h, C.c_int(i), C.c_int(j), _dp(_f64(meas)), _dp(_fcol(sqrt_info)))

    def set_values(self, poses=None, vels=None, lands=None):
        p = _f64(poses).reshape(-1) if poses is not None else None
        v = _f64(vels).reshape(-1) if vels is not None else None
        l = _f64(lands).reshape(-1) if lands is not None and self.NL else None
        self.L.gpo_set_values(self.h, _dp(p), _dp(v), _dp(l))

    def get_values(self):
        p = np.zeros((self.N, self.PS)); v = np.zeros((self.N, self.D)); l = np.zeros((self.NL, max(self.DL, 1)))
        self.L.gpo_get_values(self.h, _dp(p), _dp(v), _dp(l) if self.NL else None)
        return p, v, (l[:, :self.DL] if self.NL else np.zeros((0, self.DL)))

    def num_factors(self):
        return self.L.gpo_num_factors(self.h)

    def error(self):
        return self.L.gpo_error(self.h)

    def _split(self, flat, m, dims):
        out, o = [], 0
        for d in dims:
            if d == 0:
                break
            out.append(flat[o:o + m * d].reshape(d, m).T.copy())
            o += m * d
        return out

    def linearize_factor(self, k):
        """whitened ([A_1..A_n], b) of factor k — the GTSAM JacobianFactor payload"""
        A = np.zeros(12 * 6 * 5); b = np.zeros(12); dims = np.zeros(5, dtype=np.int32)
        m = self.L.gpo_linearize_factor(self.h, C.c_int(k), _dp(A), _dp(b), _ip(dims))
        return self._split(A, m, dims), b[:m].copy()

    def eval_factor(self, k, want_H=True):
        """unwhitened evaluateError of factor k -> (e, [H_1..H_n])"""
        H = np.zeros(12 * 6 * 5) if want_H else None
        e = np.zeros(12); dims = np.zeros(5, dtype=np.int32)
        m = self.L.gpo_eval_factor(self.h, C.c_int(k), _dp(e), _dp(H), _ip(dims))
        return e[:m].copy(), (self._split(H, m, dims) if want_H else None)

    def normal_equations_dense(self):
        n = self.N * 2 * self.D + self.NL * self.DL
        H = np.zeros((n, n)); g = np.zeros(n)
        rc = self.L.gpo_normal_equations_dense(self.h, _dp(H), _dp(g), C.c_int(n))
        assert rc == 0
        return H.T.copy(), g

    def check_step(self, delta_states, delta_lands, lam=0.0):
        """residual of a step against the oracle's own normal equations at the current values (any size):
        dict(residual = max |(J^T J + lam I) delta - J^T b|, rhs = max |J^T b|, step = max |delta|, linearized_error,
        scale = max(|J|^T |J| |delta|) + rhs: what a backward-stable solver's residual is small against)"""
        ds = _f64(delta_states).reshape(-1); dl = _f64(delta_lands).reshape(-1) if self.NL else np.zeros(1)
        assert ds.size == self.N * 2 * self.D
        out = np.zeros(5)
        self.L.gpo_check_step(self.h, _dp(ds), _dp(dl), C.c_double(lam), _dp(out))
        return dict(residual=out[0], rhs=out[1], step=out[2], linearized_error=out[3], scale=out[4] + out[1] + lam * out[2])

    def optimize(self, params=None, n_iter=0, use_lm=True):
        p = params if params is not None else default_params(use_lm)
        st = Stats()
        self.L.gpo_optimize(self.h, C.byref(p), C.c_int(n_iter), C.byref(st))
        return st


# ---- free functions (reference unit-test parity helpers)
def _call(name, *args):
    getattr(lib(), name)(*args)


def pose3(R, t):
    return np.concatenate([np.asarray(R, dtype=np.float64).T.ravel(), np.asarray(t, dtype=np.float64)])


def pose3_Rt(T):
    T = np.asarray(T)
    return T[:9].reshape(3, 3).T.copy(), T[9:12].copy()


def rot3_ypr(y, p, r):
    R = np.zeros(9); _call("gpo_rot3_ypr", C.c_double(y), C.c_double(p), C.c_double(r), _dp(R)); return R.reshape(3, 3).T.copy()


def rot3_wire(R):
    return np.asarray(R, dtype=np.float64).T.ravel().copy()


def pose3_expmap(xi):
    T = np.zeros(12); _call("gpo_pose3_expmap", _dp(_f64(xi)), _dp(T)); return T


def pose3_logmap(T):
    xi = np.zeros(6); _call("gpo_pose3_logmap", _dp(_f64(T)), _dp(xi)); return xi


def pose3_compose(A, B):
    Cc = np.zeros(12); _call("gpo_pose3_compose", _dp(_f64(A)), _dp(_f64(B)), _dp(Cc)); return Cc


def pose3_inverse(A):
    Cc = np.zeros(12); _call("gpo_pose3_inverse", _dp(_f64(A)), _dp(Cc)); return Cc


def rot3_expmap(w):
    R = np.zeros(9); _call("gpo_rot3_expmap", _dp(_f64(w)), _dp(R)); return R


def rot3_logmap(R):
    w = np.zeros(3); _call("gpo_rot3_logmap", _dp(_f64(R)), _dp(w)); return w


def pose2_expmap(xi):
    T = np.zeros(3); _call("gpo_pose2_expmap", _dp(_f64(xi)), _dp(T)); return T


def pose2_logmap(T):
    xi = np.zeros(3); _call("gpo_pose2_logmap", _dp(_f64(T)), _dp(xi)); return xi


def pose2_compose(A, B):
    Cc = np.zeros(3); _call("gpo_pose2_compose", _dp(_f64(A)), _dp(_f64(B)), _dp(Cc)); return Cc


def pose2_inverse(A):
    Cc = np.zeros(3); _call("gpo_pose2_inverse", _dp(_f64(A)), _dp(Cc)); return Cc


def so3_jacobian(which, w):
    J = np.zeros(9); _call("gpo_so3_jacobian", C.c_int(which), _dp(_f64(w)), _dp(J)); return J.reshape(3, 3).T.copy()


def se3_jacobian(which, xi):
    J = np.zeros(36); _call("gpo_se3_jacobian", C.c_int(which), _dp(_f64(xi)), _dp(J)); return J.reshape(6, 6).T.copy()


def body_centric(spatial, T1, T2, dt):
    v = np.zeros(6); _call("gpo_body_centric", C.c_int(spatial), _dp(_f64(T1)), _dp(_f64(T2)), C.c_double(dt), _dp(v)); return v


def lambda_psi(D, Qc, delta_t, tau):
    La = np.zeros(4 * D * D); Ps = np.zeros(4 * D * D)
    _call("gpo_lambda_psi", C.c_int(D), _dp(_fcol(Qc)), C.c_double(delta_t), C.c_double(tau), _dp(La), _dp(Ps))
    return La.reshape(2 * D, 2 * D).T.copy(), Ps.reshape(2 * D, 2 * D).T.copy()


def convert_vw_to_vb(v, w, pose):
    out = np.zeros(6); _call("gpo_convert_vw_to_vb", _dp(_f64(v)), _dp(_f64(w)), _dp(_f64(pose)), _dp(out)); return out


def convert_vb_to_vw(v6, pose):
    v = np.zeros(3); w = np.zeros(3); _call("gpo_convert_vb_to_vw", _dp(_f64(v6)), _dp(_f64(pose)), _dp(v), _dp(w)); return v, w


def interpolate(group, Qc, delta_t, tau, p1, v1, p2, v2, want_H=False, D=3):
    ps = POSE_STORAGE.get(group, D); d = TANGENT_DIM.get(group, D)
    if group == POSE3VW:
        ps, d = 12, 6
    out = np.zeros(ps); H = np.zeros(4 * d * d) if want_H else None
    _call("gpo_interpolate", C.c_int(group), C.c_int(d), _dp(_fcol(Qc)), C.c_double(delta_t), C.c_double(tau), _dp(_f64(p1)), _dp(_f64(v1)),
          _dp(_f64(p2)), _dp(_f64(v2)), _dp(out), _dp(H))
    if want_H:
        return out, [H[k * d * d:(k + 1) * d * d].reshape(d, d).T.copy() for k in range(4)]
    return out


def retract(group, pose, delta):
    """x (+) delta with the oracle's chart: Pose3/Rot3 Expmap, Pose2 first-order (GTSAM default), vector add"""
    if group in (POSE3, POSE3VW):
        return pose3_compose(pose, pose3_expmap(delta))
    if group == ROT3:
        R = np.asarray(pose).reshape(3, 3).T @ rot3_expmap(delta).reshape(3, 3).T
        return R.T.ravel().copy()
    if group == POSE2:
        return pose2_compose(pose, np.asarray(delta, dtype=np.float64))
    return np.asarray(pose, dtype=np.float64) + np.asarray(delta, dtype=np.float64)


def hardware_threads():
    return lib().gpo_hardware_threads()
