"""ctypes binding of the C ABI declared in include/gpb.h (one-to-one; no arithmetic happens in Python)."""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_ROOT = os.path.dirname(_HERE)
_LIB = None

GPB_POSE3, GPB_POSE2, GPB_ROT3, GPB_LINEAR, GPB_POSE3VW = 0, 1, 2, 3, 4
_PS = {GPB_POSE3: 12, GPB_POSE2: 3, GPB_ROT3: 9, GPB_LINEAR: 3, GPB_POSE3VW: 12}
_D = {GPB_POSE3: 6, GPB_POSE2: 3, GPB_ROT3: 3, GPB_LINEAR: 3, GPB_POSE3VW: 6}
_DL = {GPB_POSE3: 3, GPB_POSE2: 2, GPB_ROT3: 0, GPB_LINEAR: 2, GPB_POSE3VW: 3}

NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17", "-Xcompiler", "-fPIC", "-shared"]


def library_path():
    return os.path.join(_HERE, "libgpb.so")


def build_library(force=False, verbose=False):
    """nvcc-compile gpslam_b200/csrc/engine.cu for sm_100a into gpslam_b200/libgpb.so (in-tree, so it ships to the GPU box)."""
    src_dir = os.path.join(_HERE, "csrc")
    out = library_path()
    srcs = [os.path.join(src_dir, f) for f in os.listdir(src_dir)] + [os.path.join(_ROOT, "include", "gpb.h")]
    if not force and os.path.exists(out) and all(os.path.getmtime(out) >= os.path.getmtime(s) for s in srcs):
        return out
    cmd = ["nvcc"] + NVCC_FLAGS + (["-Xptxas", "-v"] if verbose else []) + ["-o", out, os.path.join(src_dir, "engine.cu")]
    subprocess.run(cmd, check=True)
    return out


class Params(C.Structure):
    _fields_ = [("max_iterations", C.c_int), ("rel_tol", C.c_double), ("abs_tol", C.c_double), ("err_tol", C.c_double),
                ("lambda_initial", C.c_double), ("lambda_factor", C.c_double), ("lambda_upper", C.c_double), ("lambda_lower", C.c_double),
                ("min_model_fidelity", C.c_double), ("use_lm", C.c_int)]


class Stats(C.Structure):
    _fields_ = [("iterations", C.c_int), ("error_initial", C.c_double), ("error_final", C.c_double), ("lambda_", C.c_double),
                ("linearize_ms", C.c_double), ("assemble_ms", C.c_double), ("solve_ms", C.c_double), ("update_ms", C.c_double),
                ("total_ms", C.c_double), ("status", C.c_int)]


class Sizes(C.Structure):
    _fields_ = [("hbm_bytes", C.c_double), ("linearise_bytes", C.c_double), ("fused_bytes", C.c_double), ("solve_bytes", C.c_double),
                ("n_gp", C.c_int), ("n_extra", C.c_int), ("n_rows", C.c_int), ("border_dim", C.c_int), ("levels", C.c_int)]


class GpbError(RuntimeError):
    pass


def lib():
    """Load libgpb.so.  Raises (never falls back) if the CUDA library has not been built."""
    global _LIB
    if _LIB is None:
        path = library_path()
        if not os.path.exists(path):
            raise GpbError("gpslam_b200: %s is missing — run `python -c 'import __graft_entry__ as g; g.build()'` (nvcc, sm_100a). "
                           "There is no CPU fallback." % path)
        L = C.CDLL(path)
        L.gpb_last_error.restype = C.c_char_p
        L.gpb_graph_create.restype = C.c_void_p
        L.gpb_graph_create.argtypes = [C.c_int, C.c_int, C.c_int, C.c_int]
        _LIB = L
    return _LIB


def device_count():
    return lib().gpb_device_count()


def dense_solve(A, b, lam=0.0, loff=0, force_blocked=False, device=0, force_small=False, old_tiny=False):
    """testing aid: the reduced-system solver alone (gpb_debug_dense_solve).  Default: the shared-memory blocked factorisation up to
    160 unknowns, the multi-CTA blocked Cholesky beyond; force_blocked: the multi-CTA one at any size; force_small / old_tiny: the
    per-column shared-memory kernels (plain loops / register-blocked trailing update)"""
    A = np.asarray(A, dtype=np.float64); R = A.shape[0]
    Af = np.ascontiguousarray(A.T).ravel(); bf = _f64(b); x = np.zeros(R)
    rc = lib().gpb_debug_dense_solve(C.c_int(device), C.c_int(R), _dp(Af), _dp(bf), C.c_double(lam), C.c_int(loff), C.c_int(1 if force_blocked else (2 if force_small else (3 if old_tiny else 0))), _dp(x))
    if rc != 0:
        raise GpbError(lib().gpb_last_error().decode())
    return x


def dmma_peak(device=0):
    """measured FP64 tensor-pipe peak (TFLOP/s)"""
    v = C.c_double()
    if lib().gpb_debug_dmma_peak(C.c_int(device), C.byref(v)) != 0:
        raise GpbError(lib().gpb_last_error().decode())
    return v.value


def latencies(device=0):
    """profiling aid (gpb_debug_latency): clocks per dependent operation"""
    out = np.zeros(8)
    if lib().gpb_debug_latency(C.c_int(device), _dp(out)) != 0:
        raise GpbError(lib().gpb_last_error().decode())
    return dict(zip(("dfma", "dmul", "dmma", "shfl64", "lds", "rsqrt_plus_add", "dadd", "ldg_l2"), out.tolist()))


def store_peak(mode, n_factors, device=0):
    """profiling aid (gpb_debug_store_peak): microseconds to write n_factors SE(3) [A|b] records in k_lin_gp's store pattern"""
    v = C.c_double()
    if lib().gpb_debug_store_peak(C.c_int(device), C.c_int(mode), C.c_int(n_factors), C.byref(v)) != 0:
        raise GpbError(lib().gpb_last_error().decode())
    return v.value


def interpolate_poses(group, x1, v1, x2, v2, delta_t, tau, want_H=False, device=0):
    """GaussianProcessInterpolator*::interpolatePose for n queries (gpb_interpolate_poses).  Returns poses [n x PS] and, with want_H,
    H [n, 4, D, D] (Hint1..Hint4)."""
    x1 = _f64(x1).reshape(-1, _PS[group]); n = len(x1)
    v1 = _f64(v1).reshape(n, _D[group]); x2 = _f64(x2).reshape(n, _PS[group]); v2 = _f64(v2).reshape(n, _D[group])
    dt = _f64(np.broadcast_to(np.atleast_1d(delta_t), (n,))); ta = _f64(np.broadcast_to(np.atleast_1d(tau), (n,)))
    D = _D[group]
    poses = np.zeros((n, _PS[group])); H = np.zeros((n, 4, D, D)) if want_H else None
    rc = lib().gpb_interpolate_poses(C.c_int(group), C.c_int(device), C.c_int(n), _dp(x1), _dp(v1), _dp(x2), _dp(v2), _dp(dt), _dp(ta), _dp(poses), _dp(H))
    if rc != 0:
        raise GpbError(lib().gpb_last_error().decode())
    if want_H:
        return poses, np.ascontiguousarray(H.transpose(0, 1, 3, 2))   # column-major blocks -> [row, col]
    return poses


def interpolate_velocities(x1, v1, x2, v2, delta_t, tau, want_H=False, device=0):
    """GaussianProcessInterpolatorLinear::interpolateVelocity for n queries (gpb_interpolate_velocities): vels [n x dim] and, with
    want_H, the four scalars (Lambda21, Lambda22, Psi21, Psi22) per query - H1..H4 are those multiples of the identity"""
    x1 = _f64(np.atleast_2d(x1)); n, dim = x1.shape
    v1 = _f64(v1).reshape(n, dim); x2 = _f64(x2).reshape(n, dim); v2 = _f64(v2).reshape(n, dim)
    dt = _f64(np.broadcast_to(np.atleast_1d(delta_t), (n,))); ta = _f64(np.broadcast_to(np.atleast_1d(tau), (n,)))
    out = np.zeros((n, dim)); H = np.zeros((n, 4)) if want_H else None
    rc = lib().gpb_interpolate_velocities(C.c_int(GPB_LINEAR), C.c_int(device), C.c_int(n), C.c_int(dim), _dp(x1), _dp(v1), _dp(x2), _dp(v2), _dp(dt), _dp(ta), _dp(out), _dp(H))
    if rc != 0:
        raise GpbError(lib().gpb_last_error().decode())
    return (out, H) if want_H else out


def nccl_unique_id():
    """128-byte ncclUniqueId for gpb_graph_init_nccl (rank 0 creates it, every rank passes the same bytes)"""
    buf = (C.c_ubyte * 128)()
    if lib().gpb_nccl_unique_id(buf) != 0:
        raise GpbError(lib().gpb_last_error().decode())
    return bytes(buf)


def default_params(use_lm=True):
    p = Params()
    lib().gpb_default_params(C.byref(p), C.c_int(1 if use_lm else 0))
    return p


def _dp(a):
    return a.ctypes.data_as(C.POINTER(C.c_double)) if a is not None else None


def _ip(a):
    return a.ctypes.data_as(C.POINTER(C.c_int)) if a is not None else None


def _f64(a):
    return np.ascontiguousarray(np.asarray(a, dtype=np.float64))


def _fcol(a):
    return np.ascontiguousarray(np.asarray(a, dtype=np.float64).T).ravel()


class Graph:
    """Trajectory factor graph resident on one B200.  Construction calls mirror the reference's factor constructors
    (gpslam.h:31-226); after finalize() everything lives in HBM and the calls below map one-to-one onto the C ABI."""

    def __init__(self, group, n_states, n_landmarks=0, dim=3):
        self.L = lib()
        self.group, self.N = group, n_states
        self.D, self.PS, self.DL = _D[group], _PS[group], _DL[group]
        self.NL = n_landmarks if self.DL else 0
        h = self.L.gpb_graph_create(group, dim, n_states, n_landmarks)
        if not h:
            raise GpbError(self.L.gpb_last_error().decode())
        self.h = C.c_void_p(h)
        self._pinned = []

    def __del__(self):
        if getattr(self, "h", None):
            for p in getattr(self, "_pinned", []):
                self.L.gpb_free_host(p)
            self._pinned = []
            self.L.gpb_graph_destroy(self.h)
            self.h = None

    def _ck(self, rc):
        if rc < 0:
            raise GpbError("gpb error %d: %s" % (rc, self.L.gpb_last_error().decode()))
        return rc

    # ---- construction
    def add_qc_model(self, Qc):
        return self._ck(self.L.gpb_add_qc_model(self.h, _dp(_fcol(Qc))))

    def add_gp_prior(self, i, delta_t, qc=0):
        i = np.ascontiguousarray(np.atleast_1d(i), dtype=np.int32)
        dt = _f64(np.broadcast_to(np.atleast_1d(delta_t), i.shape))
        self._ck(self.L.gpb_add_gp_prior(self.h, C.c_int(len(i)), _ip(i), _dp(dt), C.c_int(qc)))

    def add_interp_range(self, i, l, z, sigma, delta_t, tau, qc=0, body_P_sensor=None):
        i = np.ascontiguousarray(np.atleast_1d(i), dtype=np.int32)
        l = np.ascontiguousarray(np.broadcast_to(np.atleast_1d(l), i.shape), dtype=np.int32)
        b = lambda a: _f64(np.broadcast_to(np.atleast_1d(a), i.shape))
        bps = _f64(body_P_sensor) if body_P_sensor is not None else None
        self._ck(self.L.gpb_add_interp_range(self.h, C.c_int(len(i)), _ip(i), _ip(l), _dp(b(z)), _dp(b(sigma)), _dp(b(delta_t)), _dp(b(tau)), C.c_int(qc),
                                             _dp(bps)))

    def add_interp_gps(self, i, meas, sqrt_info, delta_t, tau, qc=0, body_P_sensor=None):
        """GPInterpolatedGPSFactorPose3: meas [n x 3] points, sqrt_info 3x3 upper-triangular R shared by the n factors"""
        i = np.ascontiguousarray(np.atleast_1d(i), dtype=np.int32)
        b = lambda a: _f64(np.broadcast_to(np.atleast_1d(a), i.shape))
        m = _f64(np.broadcast_to(np.asarray(meas, dtype=np.float64).reshape(-1, 3), (len(i), 3)))
        bps = _f64(body_P_sensor) if body_P_sensor is not None else None
        self._ck(self.L.gpb_add_interp_gps(self.h, C.c_int(len(i)), _ip(i), _dp(m), _dp(_fcol(sqrt_info)), _dp(b(delta_t)), _dp(b(tau)), C.c_int(qc), _dp(bps)))

    def add_interp_projection(self, i, l, meas, sqrt_info, delta_t, tau, K, qc=0, body_P_sensor=None):
        """GPInterpolatedProjectionFactorPose3<Cal3_S2>: meas [n x 2] image points, K = (fx, fy, s, u0, v0)"""
        i = np.ascontiguousarray(np.atleast_1d(i), dtype=np.int32)
        l = np.ascontiguousarray(np.broadcast_to(np.atleast_1d(l), i.shape), dtype=np.int32)
        b = lambda a: _f64(np.broadcast_to(np.atleast_1d(a), i.shape))
        m = _f64(np.broadcast_to(np.asarray(meas, dtype=np.float64).reshape(-1, 2), (len(i), 2)))
        bps = _f64(body_P_sensor) if body_P_sensor is not None else None
        self._ck(self.L.gpb_add_interp_projection(self.h, C.c_int(len(i)), _ip(i), _ip(l), _dp(m), _dp(_fcol(sqrt_info)), _dp(b(delta_t)), _dp(b(tau)), C.c_int(qc),
                                                   _dp(_f64(K)), _dp(bps)))

    def add_interp_attitude(self, i, delta_t, tau, nZ, sigma, bRef=(0, 0, 1), qc=0):
        i = np.ascontiguousarray(np.atleast_1d(i), dtype=np.int32)
        b = lambda a: _f64(np.broadcast_to(np.atleast_1d(a), i.shape))
        nz = _f64(np.broadcast_to(np.asarray(nZ, dtype=np.float64).reshape(-1, 3), (len(i), 3)))
        br = _f64(np.broadcast_to(np.asarray(bRef, dtype=np.float64).reshape(-1, 3), (len(i), 3)))
        self._ck(self.L.gpb_add_interp_attitude(self.h, C.c_int(len(i)), _ip(i), _dp(b(delta_t)), _dp(b(tau)), C.c_int(qc), _dp(nz), _dp(br), _dp(b(sigma))))

    def add_prior_pose(self, i, value, sqrt_info):
        self._ck(self.L.gpb_add_prior_pose(self.h, C.c_int(i), _dp(_f64(value)), _dp(_fcol(sqrt_info))))

    def add_prior_vel(self, i, value, sqrt_info):
        self._ck(self.L.gpb_add_prior_vel(self.h, C.c_int(i), _dp(_f64(value)), _dp(_fcol(sqrt_info))))

    def add_prior_landmark(self, l, value, sqrt_info):
        self._ck(self.L.gpb_add_prior_landmark(self.h, C.c_int(l), _dp(_f64(value)), _dp(_fcol(sqrt_info))))

    def add_between(self, i, j, meas, sqrt_info):
        self._ck(self.L.gpb_add_between(self.h, C.c_int(i), C.c_int(j), _dp(_f64(meas)), _dp(_fcol(sqrt_info))))

    def add_range_2d(self, i, l, z, sigma):
        self._ck(self.L.gpb_add_range_2d(self.h, C.c_int(i), C.c_int(l), C.c_double(z), C.c_double(sigma)))

    def add_range_bearing_2d(self, i, l, rng, bearing, sqrt_info):
        self._ck(self.L.gpb_add_range_bearing_2d(self.h, C.c_int(i), C.c_int(l), C.c_double(rng), C.c_double(bearing), _dp(_fcol(sqrt_info))))

    def add_odometry_2d(self, i, j, meas, sqrt_info):
        self._ck(self.L.gpb_add_odometry_2d(self.h, C.c_int(i), C.c_int(j), _dp(_f64(meas)), _dp(_fcol(sqrt_info))))

    def set_shard(self, rank, world, ext_left, ext_right):
        self._ck(self.L.gpb_graph_set_shard(self.h, C.c_int(rank), C.c_int(world), C.c_int(1 if ext_left else 0), C.c_int(1 if ext_right else 0)))

    def set_top_map(self, n_real, ntop_global, pinned_local, pinned_gtop):
        pl = np.ascontiguousarray(np.asarray(pinned_local, dtype=np.int32)); pg = np.ascontiguousarray(np.asarray(pinned_gtop, dtype=np.int32))
        self._ck(self.L.gpb_graph_set_top_map(self.h, C.c_int(n_real), C.c_int(ntop_global), C.c_int(len(pl)), _ip(pl), _ip(pg)))

    def set_allreduce(self, fn):
        """fn(device_ptr:int, count:int, cuda_stream:int) -> 0; in-place SUM over ranks, stream-ordered on cuda_stream (see gpb.h)"""
        CB = C.CFUNCTYPE(C.c_int, C.c_void_p, C.c_void_p, C.c_longlong, C.c_void_p)
        self._cb = CB(lambda ctx, ptr, count, stream: int(fn(ptr, count, stream or 0)))
        self._ck(self.L.gpb_set_allreduce(self.h, self._cb, None))

    def init_nccl(self, unique_id, rank, world):
        """the engine's own NCCL communicator (gpb_graph_init_nccl): collective over all ranks, after finalize()"""
        buf = (C.c_ubyte * 128).from_buffer_copy(unique_id)
        self._ck(self.L.gpb_graph_init_nccl(self.h, buf, C.c_int(rank), C.c_int(world)))

    def allreduces(self):
        return self.L.gpb_allreduces_last_optimize(self.h)

    def set_segment_length(self, level0=0, upper=0):
        self._ck(self.L.gpb_set_segment_length(self.h, C.c_int(level0), C.c_int(upper)))

    def set_values(self, poses=None, vels=None, lands=None):
        p = _f64(poses).reshape(-1) if poses is not None else None
        v = _f64(vels).reshape(-1) if vels is not None else None
        l = _f64(lands).reshape(-1) if lands is not None and self.NL else None
        self._ck(self.L.gpb_set_values(self.h, _dp(p), _dp(v), _dp(l)))

    def get_values(self, out=None):
        """out: optional (poses, vels, lands) arrays to fill in place (e.g. page-locked ones from alloc_values)"""
        if out is not None:
            p, v, l = out
            assert p.shape == (self.N, self.PS) and v.shape == (self.N, self.D) and p.flags.c_contiguous and v.flags.c_contiguous
            self._ck(self.L.gpb_get_values(self.h, _dp(p), _dp(v), _dp(l) if self.NL else None))
            return p, v, l
        p = np.zeros((self.N, self.PS)); v = np.zeros((self.N, self.D)); l = np.zeros((self.NL, max(self.DL, 1)))
        self._ck(self.L.gpb_get_values(self.h, _dp(p), _dp(v), _dp(l) if self.NL else None))
        return p, v, (l[:, :self.DL] if self.NL else np.zeros((0, self.DL)))

    def alloc_values(self):
        """page-locked (poses, vels, lands) arrays for set_values / get_values(out=...); freed with the graph"""
        out = []
        for shape in ((self.N, self.PS), (self.N, self.D), (self.NL, max(self.DL, 1))):
            n = max(int(np.prod(shape)), 1)
            ptr = C.c_void_p()
            self._ck(self.L.gpb_alloc_host(C.byref(ptr), C.c_longlong(8 * n)))
            self._pinned.append(ptr)
            out.append(np.ctypeslib.as_array(C.cast(ptr, C.POINTER(C.c_double)), shape=(n,))[:int(np.prod(shape))].reshape(shape))
        return tuple(out)

    def interpolate(self, interval, tau):
        """poses of the current estimate at time tau[k] into interval[k] (gpb_graph_interpolate): dense trajectory output"""
        iv = np.ascontiguousarray(np.atleast_1d(interval), dtype=np.int32)
        ta = _f64(np.broadcast_to(np.atleast_1d(tau), iv.shape))
        out = np.zeros((len(iv), self.PS))
        self._ck(self.L.gpb_graph_interpolate(self.h, C.c_int(len(iv)), _ip(iv), _dp(ta), _dp(out)))
        return out

    def finalize(self, device=0):
        self._ck(self.L.gpb_graph_finalize(self.h, C.c_int(device)))

    # ---- hot path
    def error(self):
        e = C.c_double()
        self._ck(self.L.gpb_error(self.h, C.byref(e)))
        return e.value

    def linearize(self):
        e = C.c_double()
        self._ck(self.L.gpb_linearize(self.h, C.byref(e)))
        return e.value

    def _split(self, flat, m, dims):
        out, o = [], 0
        for d in dims:
            if d == 0:
                break
            out.append(flat[o:o + m * d].reshape(d, m).T.copy())
            o += m * d
        return out

    def linearized_factor(self, kind, idx):
        """whitened ([A_1..A_n], b) of one factor after linearize(): kind 0 = GP prior of interval idx, 1 = idx-th other factor"""
        A = np.zeros(12 * 6 * 5); b = np.zeros(12); dims = np.zeros(5, dtype=np.int32)
        m = self._ck(self.L.gpb_get_linearized_factor(self.h, C.c_int(kind), C.c_int(idx), _dp(A), _dp(b), _ip(dims)))
        return self._split(A, m, dims), b[:m].copy()

    def normal_equations_dense(self):
        n = self.N * 2 * self.D + self.NL * self.DL
        H = np.zeros((n, n)); g = np.zeros(n)
        self._ck(self.L.gpb_get_normal_equations(self.h, _dp(H), _dp(g), C.c_int(n)))
        return H.T.copy(), g

    def solve_delta(self, lam=0.0):
        ds = np.zeros((self.N, 2 * self.D)); dl = np.zeros(max(self.NL * self.DL, 1))
        self._ck(self.L.gpb_solve_delta(self.h, C.c_double(lam), _dp(ds), _dp(dl)))
        return ds, dl[:self.NL * self.DL]

    def optimize(self, params=None, n_iter=0, use_lm=True):
        p = params if params is not None else default_params(use_lm)
        st = Stats()
        self._ck(self.L.gpb_optimize(self.h, C.byref(p), C.c_int(n_iter), C.byref(st)))
        return st

    def optimize_batch(self, ins, outs):
        """gpb_optimize_batch: ins / outs are lists of (poses, vels, lands) array triples (page-locked ones from alloc_values for
        full copy / compute overlap); one Gauss-Newton iteration per step.  Returns (stats, this rank's error after each step)."""
        K = len(ins)
        assert K == len(outs) and K >= 1
        PP = C.POINTER(C.c_double)
        def arr(trip, idx):
            a = (PP * K)()
            for k, t in enumerate(trip):
                x = t[idx]
                assert x.dtype == np.float64 and x.flags.c_contiguous
                a[k] = _dp(x)
            return a
        st = Stats(); errs = np.zeros(K)
        keep = [arr(ins, 0), arr(ins, 1), arr(ins, 2) if self.NL else None, arr(outs, 0), arr(outs, 1), arr(outs, 2) if self.NL else None]
        self._ck(self.L.gpb_optimize_batch(self.h, C.c_int(K), keep[0], keep[1], keep[2], keep[3], keep[4], keep[5], _dp(errs), C.byref(st)))
        return st, errs

    def time_stage(self, stage, reps=10):
        ms = C.c_double()
        self._ck(self.L.gpb_time_stage(self.h, C.c_int(stage), C.c_int(reps), C.byref(ms)))
        return ms.value

    def sizes(self):
        s = Sizes()
        self._ck(self.L.gpb_get_sizes(self.h, C.byref(s)))
        return s

    def launches(self):
        return self.L.gpb_kernel_launches_last_optimize(self.h)
