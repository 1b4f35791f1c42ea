"""Host-side mirror of the reference's FactorGraph / RobotPlugin interface for the
GBP hot path, forwarding to the CUDA engine through the C ABI (include/gbp_b200.h).

Method names follow the reference (crates/magics/src/planner/robot.rs and
factorgraph/factorgraph.rs): `update_topology` = update_robot_neighbours +
delete/create_interrobot_factors, `update_prior_of_horizon_state`,
`update_prior_of_current_state`, `iterate` = iterate_gbp_v2, and the four
`*_iteration` half-steps.  There is no CPU path: the library must be built
(`__graft_entry__.build()`) and a CUDA device must be present.
"""
from __future__ import annotations

import ctypes as C
import os
import weakref

import numpy as np

from .config import CConfig, GbpConfig

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None


def library_path() -> str:
    # GBP_B200_LIB selects an alternative build of the same engine (kernel tuning experiments)
    return os.environ.get("GBP_B200_LIB") or os.path.join(_HERE, "libgbp_b200.so")


def load_library() -> C.CDLL:
    """dlopen the in-tree CUDA engine; raises if it has not been built."""
    global _LIB
    if _LIB is not None:
        return _LIB
    path = library_path()
    if not os.path.exists(path):
        raise RuntimeError(
            f"{path} is missing: build the CUDA engine first (python -c 'import __graft_entry__ as g; g.build()'). "
            "magics_b200 has no CPU fallback."
        )
    lib = C.CDLL(path)
    lib.gbp_last_error.restype = C.c_char_p
    lib.gbp_world_create.restype = C.c_void_p
    lib.gbp_world_create.argtypes = [C.POINTER(CConfig), C.c_int32]
    lib.gbp_world_destroy.argtypes = [C.c_void_p]
    lib.gbp_world_read_connections.restype = C.c_int64
    lib.gbp_world_kernel_launches.restype = C.c_int64
    lib.gbp_world_kernel_launches.argtypes = [C.c_void_p]
    lib.gbp_world_num_robots.argtypes = [C.c_void_p]
    lib.gbp_host_alloc_pinned.restype = C.c_void_p
    lib.gbp_host_alloc_pinned.argtypes = [C.c_size_t]
    lib.gbp_host_free_pinned.argtypes = [C.c_void_p]
    lib.gbp_world_create_shard.restype = C.c_void_p
    lib.gbp_world_create_shard.argtypes = [C.POINTER(CConfig), C.c_int32, C.c_int32, C.c_int32, C.c_char_p]
    lib.gbp_world_create_local_shards.argtypes = [C.POINTER(CConfig), C.c_int32, C.c_int32, C.POINTER(C.c_void_p)]
    lib.gbp_world_first_global_id.restype = C.c_int64
    lib.gbp_world_first_global_id.argtypes = [C.c_void_p]
    lib.gbp_world_num_robots_global.restype = C.c_int64
    lib.gbp_world_num_robots_global.argtypes = [C.c_void_p]
    lib.gbp_world_num_ghosts.argtypes = [C.c_void_p]
    _LIB = lib
    return lib


def _p(a, t):
    return a.ctypes.data_as(C.POINTER(t)) if a is not None else None


def gbp_schedule(kind: int, internal: int, external: int):
    """GbpSchedule::schedule (gbp_schedule/src/schedules/mod.rs:60-81)."""
    lib = load_library()
    oi = np.zeros(256, np.uint8)
    oe = np.zeros(256, np.uint8)
    n = lib.gbp_schedule(C.c_int32(kind), C.c_uint8(internal), C.c_uint8(external), _p(oi, C.c_uint8), _p(oe, C.c_uint8))
    if n < 0:
        raise ValueError(lib.gbp_last_error().decode())
    return oi[:n].astype(bool), oe[:n].astype(bool)


def get_variable_timesteps(lookahead_horizon: int, lookahead_multiple: int) -> np.ndarray:
    """utils::get_variable_timesteps (utils.rs:35-75)."""
    lib = load_library()
    out = np.zeros(4096, np.uint32)
    n = lib.gbp_variable_timesteps(C.c_uint32(lookahead_horizon), C.c_uint32(lookahead_multiple), _p(out, C.c_uint32), 4096)
    if n < 0:
        raise ValueError(lib.gbp_last_error().decode())
    return out[:n].copy()


def env_to_sdf_image(env, device: int = 0) -> np.ndarray:
    """`env_to_png::env_to_sdf_image` (crates/env_to_png/src/lib.rs:149-163) on the device -> (h, w, 3) u8."""
    lib = load_library()
    h, w = env.image_shape
    out = np.empty((h, w, 3), np.uint8)
    ce, keep = env.c_struct()
    rc = lib.gbp_env_to_sdf_image(C.byref(ce), C.c_int32(device), _p(out, C.c_uint8))
    del keep
    if rc != 0:
        raise RuntimeError(f"gbp_env_to_sdf_image failed ({rc}): {lib.gbp_last_error().decode()}")
    return out


def pinned_empty(shape, dtype=np.float64) -> np.ndarray:
    """numpy array backed by page-locked host memory (gbp_host_alloc_pinned): host buffers handed to
    the upload / read-back calls DMA at PCIe speed instead of being staged by the driver.
    The block is released when the last view of the array is garbage collected."""
    lib = load_library()
    dtype = np.dtype(dtype)
    n = int(np.prod(shape))
    nbytes = max(1, n * dtype.itemsize)
    ptr = lib.gbp_host_alloc_pinned(C.c_size_t(nbytes))
    if not ptr:
        raise MemoryError("gbp_host_alloc_pinned failed: " + lib.gbp_last_error().decode())
    buf = (C.c_uint8 * nbytes).from_address(ptr)
    weakref.finalize(buf, lib.gbp_host_free_pinned, C.c_void_p(ptr))
    return np.frombuffer(buf, dtype=dtype, count=n).reshape(shape)


COMM_ID_BYTES = 128


class ReachedWhen(C.Structure):
    """gbp_reached_when_t"""
    _fields_ = [("intersects_with", C.c_int32), ("variable_index", C.c_int32), ("distance", C.c_int32),
                ("meter", C.c_float)]


def comm_unique_id() -> bytes:
    """gbp_comm_unique_id: the NCCL communicator id rank 0 hands to every rank (any transport)."""
    lib = load_library()
    buf = C.create_string_buffer(COMM_ID_BYTES)
    rc = lib.gbp_comm_unique_id(buf)
    if rc < 0:
        raise RuntimeError("gbp_comm_unique_id failed: " + lib.gbp_last_error().decode())
    return buf.raw


class World:
    """All robots' factor graphs on one B200 (device store + fused kernels), or one shard of a
    swarm partitioned over several B200s (`World.create_shard`, one process per GPU)."""

    def __init__(self, cfg: GbpConfig, device: int = 0, _handle=None):
        self._lib = load_library()
        self.cfg = cfg
        self.V = int(cfg.num_variables)
        self._c = cfg.to_c()
        self._h = _handle if _handle is not None else self._lib.gbp_world_create(C.byref(self._c), int(device))
        if not self._h:
            raise RuntimeError("gbp_world_create failed: " + self._lib.gbp_last_error().decode())

    @classmethod
    def create_shard(cls, cfg: GbpConfig, device: int, rank: int, world_size: int, comm_id: bytes | None):
        """gbp_world_create_shard: this process' shard; NCCL send/recv is the transport."""
        lib = load_library()
        c = cfg.to_c()
        h = lib.gbp_world_create_shard(C.byref(c), int(device), int(rank), int(world_size), comm_id)
        if not h:
            raise RuntimeError("gbp_world_create_shard failed: " + lib.gbp_last_error().decode())
        return cls(cfg, device, _handle=h)

    @classmethod
    def create_local_shards(cls, cfg: GbpConfig, world_size: int, device: int = 0):
        """gbp_world_create_local_shards: every shard in this process on one device."""
        lib = load_library()
        c = cfg.to_c()
        out = (C.c_void_p * world_size)()
        rc = lib.gbp_world_create_local_shards(C.byref(c), int(device), int(world_size), out)
        if rc < 0:
            raise RuntimeError("gbp_world_create_local_shards failed: " + lib.gbp_last_error().decode())
        return [cls(cfg, device, _handle=out[r]) for r in range(world_size)]

    def commit_shards(self):
        self._call("gbp_world_commit_shards")

    @property
    def first_global_id(self) -> int:
        return int(self._lib.gbp_world_first_global_id(C.c_void_p(self._h)))

    @property
    def num_robots_global(self) -> int:
        return int(self._lib.gbp_world_num_robots_global(C.c_void_p(self._h)))

    @property
    def num_ghosts(self) -> int:
        return int(self._lib.gbp_world_num_ghosts(C.c_void_p(self._h)))

    def close(self):
        if getattr(self, "_h", None):
            self._lib.gbp_world_destroy(C.c_void_p(self._h))
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _call(self, name, *args):
        rc = getattr(self._lib, name)(C.c_void_p(self._h), *args)
        if rc < 0:
            raise RuntimeError(f"{name} failed ({rc}): " + self._lib.gbp_last_error().decode())
        return rc

    @property
    def num_robots(self) -> int:
        return int(self._lib.gbp_world_num_robots(C.c_void_p(self._h)))

    @property
    def kernel_launches(self) -> int:
        return int(self._lib.gbp_world_kernel_launches(C.c_void_p(self._h)))

    # ---- construction ----------------------------------------------------
    def set_sdf(self, rgb8: np.ndarray):
        rgb8 = np.ascontiguousarray(rgb8, np.uint8)
        h, w = rgb8.shape[:2]
        self._call("gbp_world_set_sdf", _p(rgb8, C.c_uint8), C.c_int32(w), C.c_int32(h))

    def set_sdf_from_environment(self, env):
        """env_to_sdf_image on the device; the image stays there as the world's SDF."""
        ce, keep = env.c_struct()
        self._call("gbp_world_set_sdf_from_environment", C.byref(ce))
        del keep

    def add_robots(self, radii, timesteps, init_means, positions, wp_offsets, wp_xy):
        radii = np.ascontiguousarray(radii, np.float32)
        timesteps = np.ascontiguousarray(timesteps, np.uint32)
        init_means = np.ascontiguousarray(init_means, np.float64)
        positions = np.ascontiguousarray(positions, np.float32)
        wp_offsets = np.ascontiguousarray(wp_offsets, np.int32)
        wp_xy = np.ascontiguousarray(wp_xy, np.float32)
        n = radii.shape[0]
        if timesteps.shape[0] != self.V or init_means.size != n * self.V * 4:
            raise ValueError("add_robots: shape mismatch")
        self._call("gbp_world_add_robots", C.c_int32(n), _p(radii, C.c_float), _p(timesteps, C.c_uint32),
                   _p(init_means, C.c_double), _p(positions, C.c_float), _p(wp_offsets, C.c_int32),
                   _p(wp_xy, C.c_float))

    # ---- per-tick systems ------------------------------------------------
    def update_topology(self):
        self._call("gbp_world_update_topology")

    def set_comms(self, antenna_active=None, idle=None):
        a = None if antenna_active is None else np.ascontiguousarray(antenna_active, np.uint8)
        i = None if idle is None else np.ascontiguousarray(idle, np.uint8)
        self._call("gbp_world_set_comms", _p(a, C.c_uint8), _p(i, C.c_uint8))

    def set_waypoint_index(self, idx):
        idx = np.ascontiguousarray(idx, np.int32)
        self._call("gbp_world_set_waypoint_index", _p(idx, C.c_int32))

    def reached_waypoint(self, taskpoint=(0, 0, 0, 0.0), finished=(0, 0, 0, 0.0)):
        """reached_waypoint (robot.rs:2080-2176): criteria are (intersects_with, variable_index,
        distance_kind, meter) for ordinary waypoints and for the last one; returns who advanced."""
        t, f = ReachedWhen(*taskpoint), ReachedWhen(*finished)
        out = np.zeros(self.num_robots, np.uint8)
        self._call("gbp_world_reached_waypoint", C.byref(t), C.byref(f), _p(out, C.c_uint8))
        return out.astype(bool)

    def read_waypoint_index(self):
        out = np.zeros(self.num_robots, np.int32)
        self._call("gbp_world_read_waypoint_index", _p(out, C.c_int32))
        return out

    def update_robot_collisions(self):
        """update_robot_robot_collisions (planner/collisions.rs:72-143): (collisions so far, pairs colliding now)."""
        total, now = C.c_int64(0), C.c_int64(0)
        self._call("gbp_world_update_robot_collisions", C.byref(total), C.byref(now))
        return int(total.value), int(now.value)

    def read_robot_collisions(self):
        out = np.zeros(self.num_robots, np.uint32)
        self._call("gbp_world_read_robot_collisions", _p(out, C.c_uint32))
        return out

    def set_environment_colliders(self, colliders):
        """`Colliders` resource (environment.Collider list); clears the collision histories."""
        from .environment import pack_colliders
        arr, verts, _ = pack_colliders(colliders)
        self._call("gbp_world_set_environment_colliders", C.c_int32(len(colliders)), C.cast(arr, C.c_void_p),
                   C.c_int32(int(verts.shape[0]) if any(c.points for c in colliders) else 0), _p(verts, C.c_float))

    def update_environment_collisions(self):
        """update_robot_environment_collisions (planner/collisions.rs:368-431): (collisions so far, colliding now)."""
        total, now = C.c_int64(0), C.c_int64(0)
        self._call("gbp_world_update_environment_collisions", C.byref(total), C.byref(now))
        return int(total.value), int(now.value)

    def read_environment_collisions(self):
        out = np.zeros(self.num_robots, np.uint32)
        self._call("gbp_world_read_environment_collisions", _p(out, C.c_uint32))
        return out

    def set_tracking_buffers(self, capacity=10000, sample_ns=100_000_000):
        """PositionTracker::new / VelocityTracker::new (spawner.rs:627-628)."""
        self._track_capacity = int(capacity)
        self._call("gbp_world_set_tracking_buffers", C.c_int32(capacity), C.c_uint64(sample_ns))

    def track(self, delta_ns: int, elapsed_seconds: float):
        """track_positions + track_velocities for one FixedUpdate (planner/tracking.rs:117-137, 226-260)."""
        self._call("gbp_world_track", C.c_uint64(delta_ns), C.c_double(elapsed_seconds))

    def read_tracks(self):
        """Per robot, oldest sample first: (positions (k, 2) f32, velocities (m, 2) f32, timestamps (m,), measured_over
        (m,)) — `PositionTracker::positions()` and `VelocityTracker::measurements()`."""
        n, cap = self.num_robots, self._track_capacity
        npos, nvel = np.zeros(n, np.uint32), np.zeros(n, np.uint32)
        pos = np.zeros((cap, 2, n), np.float32)
        vel = np.zeros((cap, 2, n), np.float32)
        vt, vo = np.zeros((cap, n), np.float64), np.zeros((cap, n), np.float64)
        self._call("gbp_world_read_tracks", _p(npos, C.c_uint32), _p(pos, C.c_float), _p(nvel, C.c_uint32),
                   _p(vel, C.c_float), _p(vt, C.c_double), _p(vo, C.c_double))
        out = []
        for r in range(n):
            def order(k):
                k = int(k)
                return np.arange(k) if k <= cap else (np.arange(cap) + k) % cap
            ip, iv = order(npos[r]), order(nvel[r])
            out.append((pos[ip, :, r], vel[iv, :, r], vt[iv, r], vo[iv, r]))
        return out

    def update_prior_of_horizon_state(self):
        self._call("gbp_world_update_prior_of_horizon_state")

    def update_prior_of_current_state(self):
        self._call("gbp_world_update_prior_of_current_state")

    def change_prior_of_variable(self, variable_index, robots, new_means):
        robots = np.ascontiguousarray(robots, np.int32)
        new_means = np.ascontiguousarray(new_means, np.float64)
        self._call("gbp_world_change_prior_of_variable", C.c_int32(variable_index), C.c_int32(robots.shape[0]),
                   _p(robots, C.c_int32), _p(new_means, C.c_double))

    # ---- global-planner hand-off (planner/robot.rs:655-776) ----------------
    def set_tracking_path(self, robots, paths):
        """New polyline (>= 2 points each) for the listed robots: tracking path + mission waypoints, next index 1."""
        robots = np.ascontiguousarray(robots, np.int32)
        off = np.zeros(len(paths) + 1, np.int32)
        off[1:] = np.cumsum([len(p) for p in paths])
        xy = np.ascontiguousarray(np.concatenate([np.asarray(p, np.float32).reshape(-1, 2) for p in paths]), np.float32)
        self._call("gbp_world_set_tracking_path", C.c_int32(robots.size), _p(robots, C.c_int32), _p(off, C.c_int32),
                   _p(xy, C.c_float))

    def reset_variables(self, robots, means, first_last_sigma=1e30, inbetween_sigma=float("inf")):
        """FactorGraph::reset_variables for the listed robots; means (m, V, 4) f64."""
        robots = np.ascontiguousarray(robots, np.int32)
        means = np.ascontiguousarray(means, np.float64)
        assert means.size == robots.size * self.V * 4
        self._call("gbp_world_reset_variables", C.c_int32(robots.size), _p(robots, C.c_int32), _p(means, C.c_double),
                   C.c_double(first_last_sigma), C.c_double(inbetween_sigma))

    def reset_tracking_factors(self, robots):
        robots = np.ascontiguousarray(robots, np.int32)
        self._call("gbp_world_reset_tracking_factors", C.c_int32(robots.size), _p(robots, C.c_int32))

    def iterate(self):
        self._call("gbp_world_iterate")

    def iterate_schedule(self, internal, external):
        i = np.ascontiguousarray(internal, np.uint8)
        e = np.ascontiguousarray(external, np.uint8)
        self._call("gbp_world_iterate_schedule", C.c_int32(i.shape[0]), _p(i, C.c_uint8), _p(e, C.c_uint8))

    def internal_factor_iteration(self):
        self._call("gbp_world_internal_factor_iteration")

    def internal_variable_iteration(self):
        self._call("gbp_world_internal_variable_iteration")

    def external_factor_iteration(self):
        self._call("gbp_world_external_factor_iteration")

    def external_variable_iteration(self):
        self._call("gbp_world_external_variable_iteration")

    def step(self):
        self._call("gbp_world_step")

    def change_factor_enabled(self, kind, enabled):
        self._call("gbp_world_change_factor_enabled", C.c_int32(kind), C.c_uint8(int(enabled)))

    def set_safety_distance_multiplier(self, m):
        self._call("gbp_world_set_safety_distance_multiplier", C.c_float(m))

    def set_schedule(self, kind, internal, external):
        self._call("gbp_world_set_schedule", C.c_int32(kind), C.c_int32(internal), C.c_int32(external))

    # ---- read-back -------------------------------------------------------
    def read_means_into(self, out: np.ndarray):
        """Variable means of every robot into a caller-owned (n, V, 4) f64 buffer (ideally pinned)."""
        if out.dtype != np.float64 or not out.flags.c_contiguous or out.size != self.num_robots * self.V * 4:
            raise ValueError("read_means_into: need a C-contiguous f64 buffer of n*V*4 elements")
        self._call("gbp_world_read_beliefs", None, None, _p(out, C.c_double), None, None)
        return out

    def read_means_into_async(self, out: np.ndarray):
        """Start the read-back of every variable mean into a pinned (n, V, 4) buffer; the next tick may be
        launched right away, `readback_wait()` makes the buffer valid."""
        if out.dtype != np.float64 or not out.flags.c_contiguous or out.size != self.num_robots * self.V * 4:
            raise ValueError("read_means_into_async: need a C-contiguous f64 buffer of n*V*4 elements")
        self._call("gbp_world_read_beliefs_async", None, None, _p(out, C.c_double), None, None)
        return out

    def readback_wait(self):
        self._call("gbp_world_readback_wait")

    def read_beliefs(self, eta=True, lam=True, mean=True, cov=True, valid=True):
        n, V = self.num_robots, self.V
        out = {}
        a_eta = np.zeros((n, V, 4)) if eta else None
        a_lam = np.zeros((n, V, 4, 4)) if lam else None
        a_mean = np.zeros((n, V, 4)) if mean else None
        a_cov = np.zeros((n, V, 4, 4)) if cov else None
        a_valid = np.zeros((n, V), np.uint8) if valid else None
        self._call("gbp_world_read_beliefs", _p(a_eta, C.c_double), _p(a_lam, C.c_double), _p(a_mean, C.c_double),
                   _p(a_cov, C.c_double), _p(a_valid, C.c_uint8))
        if eta:
            out["eta"] = a_eta
        if lam:
            out["lam"] = a_lam
        if mean:
            out["mean"] = a_mean
        if cov:
            out["cov"] = a_cov
        if valid:
            out["valid"] = a_valid.astype(bool)
        return out

    def read_positions(self):
        xy = np.zeros((self.num_robots, 2), np.float32)
        self._call("gbp_world_read_positions", _p(xy, C.c_float))
        return xy

    def remove_robots(self, robots):
        """RobotDespawned: the robots leave the simulation (their slots stay, frozen)."""
        robots = np.ascontiguousarray(robots, np.int32)
        self._call("gbp_world_remove_robots", C.c_int32(robots.shape[0]), _p(robots, C.c_int32))

    def read_removed(self):
        out = np.zeros(self.num_robots, np.uint8)
        self._call("gbp_world_read_removed", _p(out, C.c_uint8))
        return out

    def set_message_counting(self, on: bool = True):
        """MessageCount accounting (off by default); call before add_robots."""
        self._call("gbp_world_set_message_counting", C.c_int32(int(bool(on))))

    def read_message_counts(self):
        """(n, 4) i64: sent internal / external, received internal / external per robot."""
        out = np.zeros((self.num_robots, 4), np.int64)
        self._call("gbp_world_read_message_counts", _p(out, C.c_int64))
        return out

    def export_totals(self) -> dict:
        """The per-robot totals `export.rs` writes (RobotData, export.rs:112-277) as far as the engine keeps them:
        radius, collisions.robots (planner/collisions.rs RobotRobotCollisions::get), messages sent / received
        (internal, external) when the counters are on, the mission's next waypoint index, collisions.environment
        (RobotEnvironmentCollisions::get) once colliders are set and the position / velocity histories once
        `set_tracking_buffers` has been called.  `magics_b200.export.export_data` shapes them like the reference's JSON."""
        out = {"collisions_robots": self.read_robot_collisions(), "next_waypoint": self.read_waypoint_index(),
               "removed": self.read_removed().astype(bool)}
        try:
            out["collisions_environment"] = self.read_environment_collisions()
        except RuntimeError:
            out["collisions_environment"] = None
        try:
            out["tracks"] = self.read_tracks()
        except (RuntimeError, AttributeError):
            out["tracks"] = None
        try:
            c = self.read_message_counts()
            out["messages"] = {"sent": {"internal": c[:, 0], "external": c[:, 1]},
                               "received": {"internal": c[:, 2], "external": c[:, 3]}}
        except RuntimeError:
            out["messages"] = None
        return out

    def read_tracking(self):
        """(record (n, V) i64, last_pos (n, V, 2) f32, last_value (n, V) f64) of every Tracking factor."""
        n, V = self.num_robots, self.V
        rec = np.zeros((n, V), np.int64)
        pos = np.zeros((n, V, 2), np.float32)
        val = np.zeros((n, V), np.float64)
        self._call("gbp_world_read_tracking", _p(rec, C.c_int64), _p(pos, C.c_float), _p(val, C.c_double))
        return rec, pos, val

    def read_connections(self):
        n = self.num_robots
        counts = self.node_counts()
        cap = max(1, int(counts[4]) // max(1, self.V - 1))
        off = np.zeros(n + 1, np.int64)
        nb = np.zeros(cap, np.int32)
        rn = np.zeros(cap, np.int64)
        e = self._lib.gbp_world_read_connections(C.c_void_p(self._h), _p(off, C.c_int64), _p(nb, C.c_int32),
                                                 _p(rn, C.c_int64), C.c_int64(cap))
        if e < 0:
            raise RuntimeError("gbp_world_read_connections failed: " + self._lib.gbp_last_error().decode())
        return off, nb[:e].copy(), rn[:e].copy()

    def sdf_lookup(self, xy):
        xy = np.ascontiguousarray(xy, np.float64)
        m = xy.shape[0]
        px, py, val = np.zeros(m, np.uint32), np.zeros(m, np.uint32), np.zeros(m)
        self._call("gbp_world_sdf_lookup", C.c_int32(m), _p(xy, C.c_double), _p(px, C.c_uint32), _p(py, C.c_uint32),
                   _p(val, C.c_double))
        return px, py, val

    def node_counts(self):
        out = np.zeros(5, np.int64)
        self._call("gbp_world_node_counts", _p(out, C.c_int64))
        return out

    def set_iterate_path(self, general_only: bool):
        """general_only: every robot through k_iterate (A/B tests); default: k_iterate_axis for the robots
        whose x and y chains are decoupled, decided on the device per launch."""
        self._call("gbp_world_set_iterate_path", C.c_int32(int(general_only)))

    def set_single_launch_tick(self, on: bool = True):
        """A whole `iterate` as one cooperative launch (k_tick_fused) while the swarm fits the GPU at once — the
        reference's own 10 - 50 robot scenarios; implies the general kernel for every robot."""
        self._call("gbp_world_set_iterate_path", C.c_int32(2 if on else 0))

    def read_iterate_path(self):
        """(robots currently iterated by k_iterate_axis, robots iterated by k_iterate) on this shard."""
        ax, gen = C.c_int64(0), C.c_int64(0)
        self._call("gbp_world_read_iterate_path", C.byref(ax), C.byref(gen))
        return int(ax.value), int(gen.value)

    PROFILE_KINDS = ("iterate_int", "iterate_ext", "iterate_ext_int", "topology", "priors", "halo", "iterate_general",
                     "topo_positions", "topo_search", "topo_apply", "iterate_border")

    def set_profiling(self, on: bool):
        self._call("gbp_world_set_profiling", C.c_int32(int(on)))

    def read_profile(self) -> dict:
        out = {}
        for k, name in enumerate(self.PROFILE_KINDS):
            cnt, ms = C.c_int64(0), C.c_double(0)
            try:
                self._call("gbp_world_read_profile", C.c_int32(k), C.byref(cnt), C.byref(ms))
            except RuntimeError:
                break  # an older build of the library (tuning variants) knows fewer kinds
            out[name] = {"count": int(cnt.value), "ms": float(ms.value)}
        return out

    # ---- timing helpers (CUDA events on the engine's own stream) ----------
    def sync(self):
        self._call("gbp_world_sync")

    def timer_start(self):
        self._call("gbp_world_timer_start")

    def timer_stop_ms(self) -> float:
        ms = C.c_float(0)
        self._call("gbp_world_timer_stop_ms", C.byref(ms))
        return float(ms.value)
