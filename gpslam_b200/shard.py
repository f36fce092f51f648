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


def reduced_dim(world, bs, nb, ntop=None):
    return (world - 1 if ntop is None else ntop) * bs + nb


def top_states(n_states, world, closures):
    """global list of top states (the unknowns of the reduced system besides the landmarks), in trajectory order: the last
    state of every shard but the final one, and every loop-closure endpoint"""
    s = {owned_range(n_states, r, world)[1] - 1 for r in range(world - 1)}
    for i, j in closures:
        s.add(int(i)); s.add(int(j))
    return sorted(s)


def owner_of(n_states, world, state):
    for r in range(world):
        a, b = owned_range(n_states, r, world)
        if a <= state < b:
            return r
    raise ValueError(state)


class ShardBuilder:
    """Accepts the construction calls of the GLOBAL graph (same API as gpslam_b200.Graph / the oracle) and forwards this
    rank's part, re-indexed to local state numbers, to a local graph object.

    Calls are recorded and replayed when the local graph is first needed (`.g`, `finalize`): the graph's size depends on
    the loop closures, which arrive among the other factors.  A closure (i, j) is evaluated by the rank owning min(i, j);
    when max(i, j) lives elsewhere that rank carries it as a ghost entry behind its own states.  Every closure endpoint is
    a top state on the rank that owns it (and wherever it is a ghost)."""

    def __init__(self, make_graph, group, n_states, n_landmarks, rank, world):
        self.make_graph, self.group, self.N, self.NL, self.rank, self.world = make_graph, group, n_states, n_landmarks, rank, world
        self.lo, self.hi = local_range(n_states, rank, world)
        self._calls, self._closures, self._values, self._g = [], [], None, None

    # ---- recorded construction calls
    def _rec(self, name, *args):
        if self._g is not None:
            raise RuntimeError("ShardBuilder: the local graph has already been materialised")
        self._calls.append((name, args))

    def add_qc_model(self, Qc):
        self._rec("add_qc_model", Qc)
        return sum(1 for c in self._calls if c[0] == "add_qc_model") - 1

    def add_gp_prior(self, i, delta_t, qc=0): self._rec("add_gp_prior", i, delta_t, qc)
    def add_interp_range(self, i, l, z, sigma, delta_t, tau, qc=0, body_P_sensor=None): self._rec("add_interp_range", i, l, z, sigma, delta_t, tau, qc, body_P_sensor)
    def add_interp_gps(self, i, meas, sqrt_info, delta_t, tau, qc=0, body_P_sensor=None): self._rec("add_interp_gps", i, meas, sqrt_info, delta_t, tau, qc, body_P_sensor)
    def add_interp_projection(self, i, l, meas, sqrt_info, delta_t, tau, K, qc=0, body_P_sensor=None): self._rec("add_interp_projection", i, l, meas, sqrt_info, delta_t, tau, K, qc, body_P_sensor)
    def add_interp_attitude(self, i, delta_t, tau, nZ, sigma, bRef=(0, 0, 1), qc=0): self._rec("add_interp_attitude", i, delta_t, tau, nZ, sigma, bRef, qc)
    def add_prior_pose(self, i, value, sqrt_info): self._rec("add_prior_pose", i, value, sqrt_info)
    def add_prior_vel(self, i, value, sqrt_info): self._rec("add_prior_vel", i, value, sqrt_info)
    def add_prior_landmark(self, l, value, sqrt_info): self._rec("add_prior_landmark", l, value, sqrt_info)
    def add_range_2d(self, i, l, z, sigma): self._rec("add_range_2d", i, l, z, sigma)
    def add_range_bearing_2d(self, i, l, rng, bearing, sqrt_info): self._rec("add_range_bearing_2d", i, l, rng, bearing, sqrt_info)
    def add_odometry_2d(self, i, j, meas, sqrt_info): self._rec("add_odometry_2d", i, j, meas, sqrt_info)

    def add_between(self, i, j, meas, sqrt_info):
        if abs(i - j) != 1:
            self._closures.append((int(i), int(j)))
        self._rec("add_between", i, j, meas, sqrt_info)

    def set_values(self, poses=None, vels=None, lands=None):
        if self._g is not None:
            self._apply_values(poses, vels, lands)
        else:
            self._values = (poses, vels, lands)

    # interval t (states t, t+1) is local iff lo <= t <= hi - 2
    def _own_interval(self, t):
        t = np.asarray(t)
        return (t >= self.lo) & (t <= self.hi - 2)

    def _single_state_owner(self, s):
        t = s if s <= self.N - 2 else self.N - 2
        return self.lo <= t <= self.hi - 2

    # ---- materialisation
    @property
    def g(self):
        if self._g is None:
            self._materialise()
        return self._g

    def _local(self, s):
        """local chain index of global state s: a real entry, or a ghost"""
        return s - self.lo if self.lo <= s < self.hi else self.n_real + self.ghosts.index(s)

    def _materialise(self):
        rank, world = self.rank, self.world
        self.n_real = self.hi - self.lo
        self.top = top_states(self.N, world, self._closures)
        self.ghosts = []   # global states replicated here
        for i, j in self._closures:
            if owner_of(self.N, world, min(i, j)) == rank:
                far = max(i, j)
                if not (self.lo <= far < self.hi) and far not in self.ghosts:
                    self.ghosts.append(far)
        g = self._g = self.make_graph(self.group, self.n_real + len(self.ghosts), self.NL)
        if world > 1 and hasattr(g, "set_shard"):
            g.set_shard(rank, world, rank > 0, rank < world - 1)
        # this rank's entries of the reduced system, in local chain order
        pinned = sorted([(s - self.lo, k) for k, s in enumerate(self.top) if self.lo <= s < self.hi] + [(self.n_real + q, self.top.index(s)) for q, s in enumerate(self.ghosts)])
        self.pinned_local = [p for p, _ in pinned]; self.pinned_gtop = [k for _, k in pinned]
        if self._closures and hasattr(g, "set_top_map"):
            g.set_top_map(self.n_real, len(self.top), self.pinned_local, self.pinned_gtop)
        for name, a in self._calls:
            getattr(self, "_do_" + name)(*a)
        if self._values is not None:
            self._apply_values(*self._values)

    def _do_add_qc_model(self, Qc): self._g.add_qc_model(Qc)

    def _do_add_gp_prior(self, i, delta_t, qc):
        i = np.atleast_1d(i); dt = np.broadcast_to(np.atleast_1d(delta_t), i.shape)
        m = self._own_interval(i)
        if m.any():
            self._g.add_gp_prior(i[m] - self.lo, dt[m], qc)

    def _do_add_interp_range(self, i, l, z, sigma, delta_t, tau, qc, body_P_sensor):
        i = np.atleast_1d(i); m = self._own_interval(i)
        b = lambda a: np.broadcast_to(np.atleast_1d(a), i.shape)[m]
        if m.any():
            self._g.add_interp_range(i[m] - self.lo, b(l), b(z), b(sigma), b(delta_t), b(tau), qc, body_P_sensor)

    def _do_add_interp_gps(self, i, meas, sqrt_info, delta_t, tau, qc, body_P_sensor):
        i = np.atleast_1d(i); m = self._own_interval(i)
        b = lambda a: np.broadcast_to(np.atleast_1d(a), i.shape)[m]
        if m.any():
            self._g.add_interp_gps(i[m] - self.lo, np.broadcast_to(np.asarray(meas, dtype=float).reshape(-1, 3), (len(i), 3))[m], sqrt_info, b(delta_t), b(tau), qc, body_P_sensor)

    def _do_add_interp_projection(self, i, l, meas, sqrt_info, delta_t, tau, K, qc, body_P_sensor):
        i = np.atleast_1d(i); m = self._own_interval(i)
        b = lambda a: np.broadcast_to(np.atleast_1d(a), i.shape)[m]
        if m.any():
            self._g.add_interp_projection(i[m] - self.lo, b(l), np.broadcast_to(np.asarray(meas, dtype=float).reshape(-1, 2), (len(i), 2))[m], sqrt_info, b(delta_t), b(tau), K, qc,
                                          body_P_sensor)

    def _do_add_interp_attitude(self, i, delta_t, tau, nZ, sigma, bRef, qc):
        i = np.atleast_1d(i); m = self._own_interval(i)
        b = lambda a: np.broadcast_to(np.atleast_1d(a), i.shape)[m]
        nz = np.broadcast_to(np.asarray(nZ, dtype=float).reshape(-1, 3), (len(i), 3))[m]
        br = np.broadcast_to(np.asarray(bRef, dtype=float).reshape(-1, 3), (len(i), 3))[m]
        if m.any():
            self._g.add_interp_attitude(i[m] - self.lo, b(delta_t), b(tau), nz, b(sigma), br, qc)

    def _do_add_prior_pose(self, i, value, sqrt_info):
        if self._single_state_owner(i):
            self._g.add_prior_pose(i - self.lo, value, sqrt_info)

    def _do_add_prior_vel(self, i, value, sqrt_info):
        if self._single_state_owner(i):
            self._g.add_prior_vel(i - self.lo, value, sqrt_info)

    def _do_add_prior_landmark(self, l, value, sqrt_info):
        if self.rank == 0:
            self._g.add_prior_landmark(l, value, sqrt_info)

    def _do_add_between(self, i, j, meas, sqrt_info):
        if abs(i - j) == 1:
            if self._own_interval(min(i, j)):
                self._g.add_between(i - self.lo, j - self.lo, meas, sqrt_info)
        elif owner_of(self.N, self.world, min(i, j)) == self.rank:
            self._g.add_between(self._local(i), self._local(j), meas, sqrt_info)

    def _do_add_range_2d(self, i, l, z, sigma):
        if self._single_state_owner(i):
            self._g.add_range_2d(i - self.lo, l, z, sigma)

    def _do_add_range_bearing_2d(self, i, l, rng, bearing, sqrt_info):
        if self._single_state_owner(i):
            self._g.add_range_bearing_2d(i - self.lo, l, rng, bearing, sqrt_info)

    def _do_add_odometry_2d(self, i, j, meas, sqrt_info):
        if self._own_interval(i):
            self._g.add_odometry_2d(i - self.lo, j - self.lo, meas, sqrt_info)

    def _apply_values(self, poses, vels, lands):
        sel = list(range(self.lo, self.hi)) + self.ghosts
        self._g.set_values(None if poses is None else np.asarray(poses)[sel], None if vels is None else np.asarray(vels)[sel], lands)

    def finalize(self, device=0):
        if hasattr(self.g, "finalize"):
            self.g.finalize(device)

    def __getattr__(self, name):  # everything else (optimize, linearize, get_values, ...) goes to the local graph
        if name.startswith("_"):
            raise AttributeError(name)
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


def init_engine_nccl(graph, rank, world):
    """Create the engine's own NCCL communicator on a finalized shard graph (gpb_graph_init_nccl): rank 0 draws the ncclUniqueId,
    one torch.distributed broadcast hands it to the other ranks, every rank joins.  After this the engine issues the boundary
    all-reduce itself, on its own stream, inside the captured CUDA graph of a Gauss-Newton iteration - Python is out of the loop."""
    import torch.distributed as dist
    from . import capi
    box = [capi.nccl_unique_id() if rank == 0 else None]
    dist.broadcast_object_list(box, src=0)
    graph.init_nccl(box[0], rank, world)


def reduced_index(world, rank, bs, nb, pinned_gtop=None, ntop=None):
    """global indices (in the all-reduced system [top states | landmarks]) of a rank's top-level variables, in the local order
    [its pinned chain entries..., landmarks].  Without loop closures the top states are the world-1 shard boundaries and a
    rank's pinned entries are its halo (rank > 0) and its last state (rank < world-1)."""
    if pinned_gtop is None:
        pinned_gtop = ([rank - 1] if rank > 0 else []) + ([rank] if rank < world - 1 else [])
        ntop = world - 1
    idx = []
    for k in pinned_gtop:
        idx += list(range(k * bs, (k + 1) * bs))
    idx += list(range(ntop * bs, ntop * bs + nb))
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
