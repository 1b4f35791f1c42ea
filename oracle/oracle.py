"""ctypes wrapper around oracle/_build/libgbp_oracle.so — TEST INFRASTRUCTURE.

Only tests/, __graft_entry__.smoke() and bench.py's CPU-baseline legs may import
this module; nothing in magics_b200/ does.  The wrapper mirrors the method names
of magics_b200.World so a parity test can drive both with the same code.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "_build", "libgbp_oracle.so")


class OracleConfig(C.Structure):
    # same field order as gbp_config_t (include/gbp_b200.h)
    _fields_ = [
        ("num_variables", C.c_int32),
        ("sigma_factor_dynamics", C.c_float),
        ("sigma_factor_interrobot", C.c_float),
        ("sigma_factor_obstacle", C.c_float),
        ("sigma_factor_tracking", C.c_float),
        ("safety_distance_multiplier", C.c_float),
        ("comms_radius", C.c_float),
        ("target_speed", C.c_float),
        ("delta_t", C.c_float),
        ("tracking_switch_padding", C.c_float),
        ("tracking_attraction_distance", C.c_float),
        ("enable_dynamic", C.c_uint8),
        ("enable_interrobot", C.c_uint8),
        ("enable_obstacle", C.c_uint8),
        ("enable_tracking", C.c_uint8),
        ("schedule_kind", C.c_int32),
        ("iterations_internal", C.c_int32),
        ("iterations_external", C.c_int32),
        ("world_width", C.c_double),
        ("world_height", C.c_double),
        ("strict_reference_quirks", C.c_int32),
    ]


def build(force: bool = False) -> str:
    """Compile the oracle with the committed Makefile (g++ only)."""
    src = os.path.join(_HERE, "gbp_oracle.cpp")
    if force or not os.path.exists(_LIB_PATH) or os.path.getmtime(_LIB_PATH) < os.path.getmtime(src):
        subprocess.run(["make", "-C", _HERE, "-s"] + (["-B"] if force else []), check=True)
    return _LIB_PATH


def build_fma() -> str:
    """The FMA-contracted sensitivity build (Makefile target `fma`); never the parity checker."""
    src = os.path.join(_HERE, "gbp_oracle.cpp")
    out = os.path.join(_HERE, "_build", "libgbp_oracle_fma.so")
    if not os.path.exists(out) or os.path.getmtime(out) < os.path.getmtime(src):
        subprocess.run(["make", "-C", _HERE, "-s", "fma"], check=True)
    return out


def lib_fma():
    """A second, independent instance of the library compiled with fused multiply-adds."""
    global _lib_fma
    if _lib_fma is None:
        _lib_fma = C.CDLL(build_fma())
        _lib_fma.gbpo_create.restype = C.c_void_p
        _lib_fma.gbpo_create.argtypes = [C.c_void_p]
        _lib_fma.gbpo_read_connections.restype = C.c_int64
    return _lib_fma


_lib = None
_lib_fma = None


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(_LIB_PATH):
            build()
        _lib = C.CDLL(_LIB_PATH)
        _lib.gbpo_create.restype = C.c_void_p
        _lib.gbpo_create.argtypes = [C.c_void_p]
        _lib.gbpo_read_connections.restype = C.c_int64
        assert _lib.gbpo_config_size() == C.sizeof(OracleConfig)
    return _lib


def _p(a, t):
    return a.ctypes.data_as(C.POINTER(t)) if a is not None else None


def schedule(kind: int, internal: int, external: int):
    oi = np.zeros(256, np.uint8)
    oe = np.zeros(256, np.uint8)
    n = lib().gbpo_schedule(C.c_int32(kind), C.c_uint8(internal), C.c_uint8(external),
                            _p(oi, C.c_uint8), _p(oe, C.c_uint8))
    assert n >= 0
    return oi[:n].astype(bool), oe[:n].astype(bool)


def variable_timesteps(horizon: int, multiple: int) -> np.ndarray:
    out = np.zeros(4096, np.uint32)
    n = lib().gbpo_variable_timesteps(C.c_uint32(horizon), C.c_uint32(multiple), _p(out, C.c_uint32), 4096)
    assert n >= 0
    return out[:n].copy()


def marginalise(eta: np.ndarray, lam: np.ndarray, marg_idx: int):
    """marginalise_factor_distance; returns None for Message::empty()."""
    eta = np.ascontiguousarray(eta, np.float64)
    lam = np.ascontiguousarray(lam, np.float64)
    n = eta.shape[0]
    oeta, olam, omu = np.zeros(n), np.zeros((n, n)), np.zeros(n)
    k = lib().gbpo_marginalise(n, _p(eta, C.c_double), _p(lam, C.c_double), marg_idx,
                               _p(oeta, C.c_double), _p(olam, C.c_double), _p(omu, C.c_double))
    if k == 0:
        return None
    return oeta[:k].copy(), olam.reshape(-1)[: k * k].reshape(k, k).copy(), omu[:k].copy()


def extract_blocks(lam8: np.ndarray, marg_idx: int):
    lam8 = np.ascontiguousarray(lam8, np.float64)
    aa, ab, ba, bb = (np.zeros((4, 4)) for _ in range(4))
    lib().gbpo_extract_blocks(_p(lam8, C.c_double), marg_idx, _p(aa, C.c_double), _p(ab, C.c_double),
                              _p(ba, C.c_double), _p(bb, C.c_double))
    return aa, ab, ba, bb


def inv4(m: np.ndarray):
    m = np.ascontiguousarray(m, np.float64)
    out = np.zeros((4, 4))
    ok = lib().gbpo_inv4(_p(m, C.c_double), _p(out, C.c_double))
    return out if ok else None


def env_to_sdf_image(env) -> np.ndarray:
    """CPU restatement of env_to_png::env_to_sdf_image (crates/env_to_png/src/lib.rs:149-163) -> (h, w, 3) u8.
    `env` is a magics_b200.environment.Environment (plain data)."""
    h, w = env.image_shape
    out = np.empty((h, w, 3), np.uint8)
    codes = env.tile_codes()
    obs, pts = env.c_obstacles()  # plain input data, same struct layout as the oracle's `envpng::Obstacle`
    rc = lib().gbpo_env_to_sdf_image(env.nrows, env.ncols, _p(codes, C.c_uint32), C.c_float(env.tile_size),
                                     C.c_float(env.path_width), C.c_uint32(env.resolution), C.c_float(env.expansion),
                                     C.c_float(env.blur), len(env.obstacles), C.cast(obs, C.c_void_p),
                                     _p(pts, C.c_double), _p(out, C.c_uint8))
    if rc != 0:
        raise RuntimeError(f"gbpo_env_to_sdf_image failed ({rc})")
    return out


class OracleWorld:
    """Same surface as magics_b200.World, executed by the CPU restatement."""

    def __init__(self, cfg, threads: int = 1, fma: bool = False):
        self._cfg = OracleConfig(**{f[0]: getattr(cfg, f[0]) for f in OracleConfig._fields_})
        self._lib = lib_fma() if fma else lib()
        self._h = self._lib.gbpo_create(C.byref(self._cfg))
        self.V = int(cfg.num_variables)
        self._lib.gbpo_set_threads(C.c_void_p(self._h), threads)

    def close(self):
        if self._h:
            self._lib.gbpo_destroy(C.c_void_p(self._h))
            self._h = None

    def __del__(self):
        self.close()

    def _call(self, name, *args):
        rc = getattr(self._lib, name)(C.c_void_p(self._h), *args)
        if rc < 0:
            raise RuntimeError(f"oracle {name} failed: {rc}")
        return rc

    @property
    def num_robots(self):
        return self._lib.gbpo_num_robots(C.c_void_p(self._h))

    def set_threads(self, t):
        self._lib.gbpo_set_threads(C.c_void_p(self._h), int(t))

    def set_sdf(self, rgb8: np.ndarray):
        rgb8 = np.ascontiguousarray(rgb8, np.uint8)
        h, w = rgb8.shape[:2]
        self._call("gbpo_set_sdf", _p(rgb8, C.c_uint8), w, h)

    def add_robots(self, radii, timesteps, init_means, positions, wp_offsets, wp_xy):
        radii = np.ascontiguousarray(radii, np.float32)
        timesteps = np.ascontiguousarray(timesteps, np.uint32)
        init_means = np.ascontiguousarray(init_means, np.float64)
        positions = np.ascontiguousarray(positions, np.float32)
        wp_offsets = np.ascontiguousarray(wp_offsets, np.int32)
        wp_xy = np.ascontiguousarray(wp_xy, np.float32)
        n = radii.shape[0]
        assert timesteps.shape[0] == self.V and init_means.size == n * self.V * 4
        self._call("gbpo_add_robots", n, _p(radii, C.c_float), _p(timesteps, C.c_uint32),
                   _p(init_means, C.c_double), _p(positions, C.c_float), _p(wp_offsets, C.c_int32),
                   _p(wp_xy, C.c_float))

    def update_topology(self):
        self._call("gbpo_update_topology")

    def set_comms(self, antenna_active=None, idle=None):
        a = None if antenna_active is None else np.ascontiguousarray(antenna_active, np.uint8)
        i = None if idle is None else np.ascontiguousarray(idle, np.uint8)
        self._call("gbpo_set_comms", _p(a, C.c_uint8), _p(i, C.c_uint8))

    def remove_robots(self, robots):
        robots = np.ascontiguousarray(robots, np.int32)
        self._call("gbpo_remove_robots", int(robots.shape[0]), _p(robots, C.c_int32))

    def set_environment_colliders(self, colliders):
        from magics_b200.environment import pack_colliders  # plain data packing, no engine involved
        _, verts, rows = pack_colliders(colliders)
        self._call("gbpo_set_environment_colliders", len(colliders), _p(rows, C.c_float), int(verts.shape[0]),
                   _p(verts, C.c_float))

    def update_environment_collisions(self):
        total, now = C.c_int64(0), C.c_int64(0)
        self._env_hits = np.zeros(self.num_robots, np.uint32)
        self._call("gbpo_update_environment_collisions", C.byref(total), C.byref(now), _p(self._env_hits, C.c_uint32))
        return int(total.value), int(now.value)

    def read_environment_collisions(self):
        return self._env_hits

    def set_tracking_buffers(self, capacity=10000, sample_ns=100_000_000):
        self._track_capacity = int(capacity)
        self._call("gbpo_set_tracking_buffers", int(capacity), C.c_uint64(sample_ns))

    def track(self, delta_ns: int, elapsed_seconds: float):
        self._call("gbpo_track", C.c_uint64(delta_ns), C.c_double(elapsed_seconds))

    def read_tracks(self):
        cap, out = self._track_capacity, []
        for r in range(self.num_robots):
            pos, vel = np.zeros((cap, 2), np.float32), np.zeros((cap, 2), np.float32)
            vt, vo = np.zeros(cap, np.float64), np.zeros(cap, np.float64)
            nv = C.c_int(0)
            k = self._call("gbpo_read_track", r, _p(pos, C.c_float), C.byref(nv), _p(vel, C.c_float), _p(vt, C.c_double),
                           _p(vo, C.c_double))
            out.append((pos[:k], vel[:nv.value], vt[:nv.value], vo[:nv.value]))
        return out

    def set_waypoint_index(self, idx):
        idx = np.ascontiguousarray(idx, np.int32)
        self._call("gbpo_set_waypoint_index", _p(idx, C.c_int32))

    def reached_waypoint(self, taskpoint=(0, 0, 0, 0.0), finished=(0, 0, 0, 0.0)):
        """(intersects_with, variable_index, distance_kind, meter) for ordinary / last waypoints."""
        crit = np.array([taskpoint[0], taskpoint[1], taskpoint[2], finished[0], finished[1], finished[2]], np.int32)
        meters = np.array([taskpoint[3], finished[3]], np.float32)
        out = np.zeros(self.num_robots, np.uint8)
        self._call("gbpo_reached_waypoint", _p(crit, C.c_int32), _p(meters, C.c_float), _p(out, C.c_uint8))
        return out.astype(bool)

    def update_robot_collisions(self):
        total, now = C.c_int64(0), C.c_int64(0)
        self._hits = np.zeros(self.num_robots, np.uint32)
        self._call("gbpo_update_robot_collisions", C.byref(total), C.byref(now), _p(self._hits, C.c_uint32))
        return int(total.value), int(now.value)

    def read_robot_collisions(self):
        return self._hits.copy()

    def read_waypoint_index(self):
        out = np.zeros(self.num_robots, np.int32)
        self._call("gbpo_read_waypoint_index", _p(out, C.c_int32))
        return out

    def read_removed(self):
        out = np.zeros(self.num_robots, np.uint8)
        self._call("gbpo_read_removed", _p(out, C.c_uint8))
        return out

    def read_collision_events(self, kind: int):
        """Every Hit so far with its Aabb intersection (CollisionHistory::aabbs, collisions.rs:463-470, :700-716):
        kind 0 robot-robot -> pairs (r, c), r < c; kind 1 robot-environment -> (robot, collider index)."""
        self._lib.gbpo_read_collision_events.restype = C.c_int64
        n = self._lib.gbpo_read_collision_events(C.c_void_p(self._h), C.c_int(kind), C.c_int64(0), None, None)
        pairs, aabbs = np.zeros((max(n, 1), 2), np.int32), np.zeros((max(n, 1), 4), np.float32)
        self._lib.gbpo_read_collision_events(C.c_void_p(self._h), C.c_int(kind), C.c_int64(n), _p(pairs, C.c_int32),
                                             _p(aabbs, C.c_float))
        return pairs[:n], aabbs[:n]

    def update_prior_of_horizon_state(self):
        self._call("gbpo_update_prior_of_horizon_state")

    def update_prior_of_current_state(self):
        self._call("gbpo_update_prior_of_current_state")

    def change_prior_of_variable(self, variable_index, robots, new_means):
        robots = np.ascontiguousarray(robots, np.int32)
        new_means = np.ascontiguousarray(new_means, np.float64)
        self._call("gbpo_change_prior_of_variable", int(variable_index), robots.shape[0],
                   _p(robots, C.c_int32), _p(new_means, C.c_double))

    def set_tracking_path(self, robots, paths):
        robots = np.ascontiguousarray(robots, np.int32)
        off = np.zeros(len(paths) + 1, np.int32)
        off[1:] = np.cumsum([len(p) for p in paths])
        xy = np.ascontiguousarray(np.concatenate([np.asarray(p, np.float32).reshape(-1, 2) for p in paths]), np.float32)
        self._call("gbpo_set_tracking_path", int(robots.size), _p(robots, C.c_int32), _p(off, C.c_int32),
                   _p(xy, C.c_float))

    def reset_variables(self, robots, means, first_last_sigma=1e30, inbetween_sigma=float("inf")):
        robots = np.ascontiguousarray(robots, np.int32)
        means = np.ascontiguousarray(means, np.float64)
        self._call("gbpo_reset_variables", int(robots.size), _p(robots, C.c_int32), _p(means, C.c_double),
                   C.c_double(first_last_sigma), C.c_double(inbetween_sigma))

    def reset_tracking_factors(self, robots):
        robots = np.ascontiguousarray(robots, np.int32)
        self._call("gbpo_reset_tracking_factors", int(robots.size), _p(robots, C.c_int32))

    def iterate(self):
        self._call("gbpo_iterate")

    def iterate_schedule(self, internal, external):
        i = np.ascontiguousarray(internal, np.uint8)
        e = np.ascontiguousarray(external, np.uint8)
        self._call("gbpo_iterate_schedule", i.shape[0], _p(i, C.c_uint8), _p(e, C.c_uint8))

    def internal_factor_iteration(self):
        self._call("gbpo_internal_factor_iteration")

    def internal_variable_iteration(self):
        self._call("gbpo_internal_variable_iteration")

    def external_factor_iteration(self):
        self._call("gbpo_external_factor_iteration")

    def external_variable_iteration(self):
        self._call("gbpo_external_variable_iteration")

    def step(self):
        self._call("gbpo_step")

    def change_factor_enabled(self, kind, enabled):
        self._call("gbpo_change_factor_enabled", int(kind), C.c_uint8(int(enabled)))

    def set_safety_distance_multiplier(self, m):
        self._call("gbpo_set_safety_distance_multiplier", C.c_float(m))

    def set_schedule(self, kind, internal, external):
        self._call("gbpo_set_schedule", int(kind), int(internal), int(external))

    def read_beliefs(self):
        n, V = self.num_robots, self.V
        eta = np.zeros((n, V, 4))
        lam = np.zeros((n, V, 4, 4))
        mean = np.zeros((n, V, 4))
        cov = np.zeros((n, V, 4, 4))
        valid = np.zeros((n, V), np.uint8)
        self._call("gbpo_read_beliefs", _p(eta, C.c_double), _p(lam, C.c_double), _p(mean, C.c_double),
                   _p(cov, C.c_double), _p(valid, C.c_uint8))
        return dict(eta=eta, lam=lam, mean=mean, cov=cov, valid=valid.astype(bool))

    def read_positions(self):
        xy = np.zeros((self.num_robots, 2), np.float32)
        self._call("gbpo_read_positions", _p(xy, C.c_float))
        return xy

    def read_connections(self, capacity=None):
        n = self.num_robots
        cap = capacity or max(1, n * 64)
        while True:
            off = np.zeros(n + 1, np.int64)
            nb = np.zeros(cap, np.int32)
            rn = np.zeros(cap, np.int64)
            e = self._lib.gbpo_read_connections(C.c_void_p(self._h), _p(off, C.c_int64), _p(nb, C.c_int32),
                                                _p(rn, C.c_int64), C.c_int64(cap))
            if e >= 0:
                return off, nb[:e].copy(), rn[:e].copy()
            cap *= 4

    def sdf_lookup(self, xy):
        xy = np.ascontiguousarray(xy, np.float64)
        m = xy.shape[0]
        px, py, val = np.zeros(m, np.uint32), np.zeros(m, np.uint32), np.zeros(m)
        self._call("gbpo_sdf_lookup", m, _p(xy, C.c_double), _p(px, C.c_uint32), _p(py, C.c_uint32),
                   _p(val, C.c_double))
        return px, py, val

    def node_counts(self):
        out = np.zeros(5, np.int64)
        self._call("gbpo_node_counts", _p(out, C.c_int64))
        return out

    def read_message_counts(self):
        """(n, 4) i64: messages sent internal / external, received internal / external per robot (FactorGraph::
        messages_sent / messages_received)."""
        out = np.zeros((self.num_robots, 4), np.int64)
        self._call("gbpo_read_message_counts", _p(out, C.c_int64))
        return out

    def read_mirror_message(self, robot, var, from_robot):
        """(eta (4,), lam (4, 4)) of the message variable `var` of `robot` holds from the InterRobot factor owned by
        `from_robot`; None when the slot is absent or holds Message::empty()."""
        eta, lam = np.zeros(4), np.zeros((4, 4))
        ok = self._call("gbpo_read_mirror_message", int(robot), int(var), int(from_robot), _p(eta, C.c_double),
                        _p(lam, C.c_double))
        return (eta, lam) if ok else None

    def has_mirror_slot(self, robot, var, from_robot) -> bool:
        """whether the variable's inbox has an entry keyed by a factor of `from_robot` at all (Empty or not)."""
        return bool(self._call("gbpo_has_mirror_slot", int(robot), int(var), int(from_robot)))

    def read_tracking(self, robot, var):
        rec = C.c_int64(0)
        pos = np.zeros(2, np.float32)
        val = C.c_double(0)
        ok = self._call("gbpo_read_tracking", int(robot), int(var), C.byref(rec), _p(pos, C.c_float), C.byref(val))
        return (rec.value, pos, val.value) if ok else None
