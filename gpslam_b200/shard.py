"""Trajectory sharding across the GPUs of one node: host-side partitioning logic.

The chain is cut into contiguous segments, one per rank.  Rank k owns states [a_k, b_k); its local chain is
[a_k - 1 (halo copy of the left neighbour's last state, if k > 0), ..., b_k - 1].  Its last state (k < P-1) and the halo are
separators of the global reduced system.  A factor attached to interval (t, t+1) belongs to the one rank whose local chain
contains both states, so every factor is evaluated exactly once.  Landmarks are replicated; landmark priors live on rank 0.
"""
import numpy as np


def owned_range(n_states, rank, world):
    base, rem = divmod(n_states, world)
    a = rank * base + min(rank, rem)
    return a, a + base + (1 if rank < rem else 0)


def local_range(n_states, rank, world):
    a, b = owned_range(n_states, rank, world)
    return a - (1 if rank > 0 else 0), b


def reduced_dim(world, bs, nb):
    return (world - 1) * bs + nb


class ShardBuilder:
    """Accepts the construction calls of the GLOBAL graph (same API as gpslam_b200.Graph / the oracle) and forwards this
    rank's part, re-indexed to local state numbers, to a local graph object."""

    def __init__(self, make_graph, group, n_states, n_landmarks, rank, world):
        self.N, self.rank, self.world = n_states, rank, world
        self.lo, self.hi = local_range(n_states, rank, world)
        self.g = make_graph(group, self.hi - self.lo, n_landmarks)
        if world > 1 and hasattr(self.g, "set_shard"):
            self.g.set_shard(rank, world, rank > 0, rank < world - 1)

    # interval t (states t, t+1) is local iff lo <= t <= hi - 2
    def _own_interval(self, t):
        t = np.asarray(t)
        return (t >= self.lo) & (t <= self.hi - 2)

    def _single_state_owner(self, s):
        t = s if s <= self.N - 2 else self.N - 2
        return self.lo <= t <= self.hi - 2

    def add_qc_model(self, Qc):
        return self.g.add_qc_model(Qc)

    def add_gp_prior(self, i, delta_t, qc=0):
        i = np.atleast_1d(i); dt = np.broadcast_to(np.atleast_1d(delta_t), i.shape)
        m = self._own_interval(i)
        if m.any():
            self.g.add_gp_prior(i[m] - self.lo, dt[m], qc)

    def add_interp_range(self, i, l, z, sigma, delta_t, tau, qc=0, body_P_sensor=None):
        i = np.atleast_1d(i); m = self._own_interval(i)
        b = lambda a: np.broadcast_to(np.atleast_1d(a), i.shape)[m]
        if m.any():
            self.g.add_interp_range(i[m] - self.lo, b(l), b(z), b(sigma), b(delta_t), b(tau), qc, body_P_sensor)

    def add_interp_attitude(self, i, delta_t, tau, nZ, sigma, bRef=(0, 0, 1), qc=0):
        i = np.atleast_1d(i); m = self._own_interval(i)
        b = lambda a: np.broadcast_to(np.atleast_1d(a), i.shape)[m]
        nz = np.broadcast_to(np.asarray(nZ, dtype=float).reshape(-1, 3), (len(i), 3))[m]
        br = np.broadcast_to(np.asarray(bRef, dtype=float).reshape(-1, 3), (len(i), 3))[m]
        if m.any():
            self.g.add_interp_attitude(i[m] - self.lo, b(delta_t), b(tau), nz, b(sigma), br, qc)

    def add_prior_pose(self, i, value, sqrt_info):
        if self._single_state_owner(i):
            self.g.add_prior_pose(i - self.lo, value, sqrt_info)

    def add_prior_vel(self, i, value, sqrt_info):
        if self._single_state_owner(i):
            self.g.add_prior_vel(i - self.lo, value, sqrt_info)

    def add_prior_landmark(self, l, value, sqrt_info):
        if self.rank == 0:
            self.g.add_prior_landmark(l, value, sqrt_info)

    def add_between(self, i, j, meas, sqrt_info):
        if abs(i - j) == 1 and self._own_interval(min(i, j)):
            self.g.add_between(i - self.lo, j - self.lo, meas, sqrt_info)
        elif abs(i - j) != 1:
            raise NotImplementedError("loop closures are not supported by the sharded build")

    def add_range_2d(self, i, l, z, sigma):
        if self._single_state_owner(i):
            self.g.add_range_2d(i - self.lo, l, z, sigma)

    def add_range_bearing_2d(self, i, l, rng, bearing, sqrt_info):
        if self._single_state_owner(i):
            self.g.add_range_bearing_2d(i - self.lo, l, rng, bearing, sqrt_info)

    def add_odometry_2d(self, i, j, meas, sqrt_info):
        if self._own_interval(i):
            self.g.add_odometry_2d(i - self.lo, j - self.lo, meas, sqrt_info)

    def set_values(self, poses=None, vels=None, lands=None):
        self.g.set_values(None if poses is None else np.asarray(poses)[self.lo:self.hi], None if vels is None else np.asarray(vels)[self.lo:self.hi], lands)

    def finalize(self, device=0):
        if hasattr(self.g, "finalize"):
            self.g.finalize(device)

    def __getattr__(self, name):  # everything else (optimize, linearize, get_values, ...) goes to the local graph
        return getattr(self.g, name)


def torch_allreduce(device):
    """gpb_allreduce_fn over torch.distributed (NCCL over NVLink): wraps the engine's device buffer without a copy and enqueues
    the collective in order on the engine's own stream - no host synchronisation, so the GN loop keeps running ahead"""
    import torch
    import torch.distributed as dist

    class _Ptr:
        def __init__(self, ptr, n):
            self.__cuda_array_interface__ = {"shape": (n,), "typestr": "<f8", "data": (ptr, False), "version": 2}

    dev = torch.device("cuda", device)

    cache = {}

    def fn(ptr, count, stream):
        key = (ptr, count, stream)
        ent = cache.get(key)
        if ent is None:
            ent = cache[key] = (torch.as_tensor(_Ptr(ptr, count), device=dev), torch.cuda.ExternalStream(stream, device=dev))
        t, ext = ent
        with torch.cuda.stream(ext):
            dist.all_reduce(t)
        return 0
    return fn


def reduced_index(world, rank, bs, nb):
    """global indices (in the all-reduced system [sep_0..sep_{P-2} | landmarks]) of a rank's top-level variables, in the local
    order [left external separator (if rank > 0), right external separator (if rank < P-1), landmarks]"""
    idx = []
    if rank > 0:
        idx += list(range((rank - 1) * bs, rank * bs))
    if rank < world - 1:
        idx += list(range(rank * bs, (rank + 1) * bs))
    idx += list(range((world - 1) * bs, (world - 1) * bs + nb))
    return np.asarray(idx, dtype=np.int64)


class LocalAllreduce:
    """In-process all-reduce for several shard graphs living on ONE GPU (one Python thread per shard): used to exercise the
    sharded code path (pack -> all-reduce -> redundant top solve -> back-substitution) where only a single GPU is available."""

    def __init__(self, world, lib):
        import threading
        self.world, self.lib = world, lib
        self.bar = threading.Barrier(world)
        self.slots = [None] * world
        self.sum = None

    def make(self, rank):
        import ctypes as C

        def fn(ptr, count, stream):
            self.lib.gpb_stream_synchronize(C.c_void_p(stream))
            self.slots[rank] = (ptr, count)
            self.bar.wait()
            if rank == 0:
                acc = np.zeros(count)
                tmp = np.zeros(count)
                for p, n in self.slots:
                    assert n == count
                    self.lib.gpb_memcpy(tmp.ctypes.data_as(C.c_void_p), C.c_void_p(p), C.c_longlong(8 * count), C.c_int(2))
                    acc += tmp
                self.sum = acc
            self.bar.wait()
            self.lib.gpb_memcpy(C.c_void_p(ptr), self.sum.ctypes.data_as(C.c_void_p), C.c_longlong(8 * count), C.c_int(1))
            self.bar.wait()
            return 0
        return fn
