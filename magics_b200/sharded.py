"""Host-side helpers for a swarm partitioned over several shards.

`partition` decides which contiguous global-id range each rank owns; `LocalShards`
drives `world_size` shards that all live in this process on ONE device through the
same code path a multi-GPU run takes (ghost slots, cross-shard robot_number exchange,
per-sub-step halo), with device-to-device copies instead of NCCL.  It mirrors the
`World` interface over the union of the shards, so a parity test can run one scenario
against a plain `World`, a `LocalShards` and the oracle and compare all three.
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from .config import GbpConfig
from .world import World


def partition(n: int, world_size: int) -> np.ndarray:
    """Contiguous id ranges in rank order: rank q owns [b[q], b[q+1]); sizes differ by at most one."""
    base, rem = divmod(int(n), int(world_size))
    sizes = np.full(world_size, base, np.int64)
    sizes[:rem] += 1
    return np.concatenate([[0], np.cumsum(sizes)]).astype(np.int64)


_M1 = np.uint64(0x9E3779B97F4A7C15)
_M2 = np.uint64(0xBF58476D1CE4E5B9)
_M3 = np.uint64(0x94D049BB133111EB)


def _mix64(x: np.ndarray) -> np.ndarray:
    """splitmix64 finaliser, element-wise on uint64 (wrapping arithmetic)."""
    with np.errstate(over="ignore"):
        x = (x ^ (x >> np.uint64(30))) * _M2
        x = (x ^ (x >> np.uint64(27))) * _M3
        return x ^ (x >> np.uint64(31))


def state_hash(first_global_id: int, means: np.ndarray, offsets: np.ndarray, neighbours: np.ndarray,
               robot_number: np.ndarray) -> tuple[int, int]:
    """Partition-independent 64-bit digests of a shard's state, to be SUMMED modulo 2**64 over the shards:
    (means, connectivity).  `means`: (n, V, 4) f64 — every bit of every variable mean enters, keyed by the
    robot's global id and the component index; connectivity: every directed pair (robot, neighbour,
    robot_number) in global ids.  A swarm partitioned over 1, 2, 4 or 8 GPUs gives the same two sums
    iff all means, the whole InterRobot edge set and every robot_number are bit-identical."""
    with np.errstate(over="ignore"):
        m = np.ascontiguousarray(means, np.float64)
        n = m.shape[0]
        bits = m.reshape(n, m.size // n if n else 0).view(np.uint64)
        gid = (np.arange(n, dtype=np.uint64) + np.uint64(first_global_id))[:, None]
        comp = np.arange(bits.shape[1], dtype=np.uint64)[None, :]
        h_means = int(np.sum(_mix64(bits ^ _mix64(gid * _M1 + comp)), dtype=np.uint64)) if n else 0
        off = np.asarray(offsets, np.int64)
        deg = np.diff(off)
        src = np.repeat(np.arange(n, dtype=np.uint64) + np.uint64(first_global_id), deg)
        nb = np.asarray(neighbours).astype(np.uint64)
        rn = np.asarray(robot_number).astype(np.uint64)
        h_conn = int(np.sum(_mix64(_mix64(src * _M1 + nb) ^ (rn * _M2)), dtype=np.uint64)) if src.size else 0
    return h_means, h_conn


class LocalShards:
    """`world_size` shards of one swarm in this process on one device (gbp_world_create_local_shards)."""

    def __init__(self, cfg: GbpConfig, world_size: int, device: int = 0, bounds=None):
        self.cfg = cfg
        self.V = int(cfg.num_variables)
        self.ws = int(world_size)
        self.shards = World.create_local_shards(cfg, world_size, device)
        self.bounds = None if bounds is None else np.asarray(bounds, np.int64)

    def close(self):
        for w in self.shards:
            w.close()
        self.shards = []

    # ---- construction ------------------------------------------------------------------
    def set_sdf(self, rgb8):
        for w in self.shards:
            w.set_sdf(rgb8)

    def add_robots(self, radii, timesteps, init_means, positions, wp_offsets, wp_xy):
        n = int(np.asarray(radii).shape[0])
        if getattr(self, "_committed", False):
            # robots spawned later take the ids above every existing one: they join the last shard, and the group
            # publishes the new total (gbp_world_commit_shards again)
            self.shards[-1].add_robots(radii, timesteps, init_means, positions, wp_offsets, wp_xy)
            self.bounds = self.bounds.copy()
            self.bounds[-1] += n
            self.shards[0].commit_shards()
            return
        self._committed = True
        if self.bounds is None:
            self.bounds = partition(n, self.ws)
        b = self.bounds
        if b.shape[0] != self.ws + 1 or b[0] != 0 or b[-1] != n or np.any(np.diff(b) < 0):
            raise ValueError("bounds must be ws+1 non-decreasing indices from 0 to n")
        init_means = np.asarray(init_means).reshape(n, self.V, 4)
        positions = np.asarray(positions).reshape(n, 2)
        wp_offsets = np.asarray(wp_offsets, np.int64)
        wp_xy = np.asarray(wp_xy).reshape(-1, 2)
        for q, w in enumerate(self.shards):
            lo, hi = int(b[q]), int(b[q + 1])
            if hi > lo:
                a, e = int(wp_offsets[lo]), int(wp_offsets[hi])
                w.add_robots(np.asarray(radii)[lo:hi], timesteps, init_means[lo:hi], positions[lo:hi],
                             (wp_offsets[lo:hi + 1] - a).astype(np.int32), wp_xy[a:e])
        self.shards[0].commit_shards()

    @property
    def num_robots(self) -> int:
        return sum(w.num_robots for w in self.shards)

    @property
    def kernel_launches(self) -> int:
        return sum(w.kernel_launches for w in self.shards)

    def _split(self, a, dtype):
        if a is None:
            return [None] * self.ws
        a = np.ascontiguousarray(a, dtype)
        return [a[int(self.bounds[q]):int(self.bounds[q + 1])] for q in range(self.ws)]

    # ---- per-tick systems: collective calls go through shard 0 and run for the whole group ----
    def update_topology(self):
        self.shards[0].update_topology()

    def set_comms(self, antenna_active=None, idle=None):
        for w, a, i in zip(self.shards, self._split(antenna_active, np.uint8), self._split(idle, np.uint8)):
            w.set_comms(a, i)

    def set_waypoint_index(self, idx):
        for w, a in zip(self.shards, self._split(idx, np.int32)):
            if w.num_robots:
                w.set_waypoint_index(a)

    def reached_waypoint(self, taskpoint=(0, 0, 0, 0.0), finished=(0, 0, 0, 0.0)):
        return np.concatenate([w.reached_waypoint(taskpoint, finished) for w in self.shards])

    def read_waypoint_index(self):
        return np.concatenate([w.read_waypoint_index() for w in self.shards])

    def update_robot_collisions(self):
        self.shards[0]._call("gbp_world_update_robot_collisions", None, None)  # collective, no sync
        total = now = 0
        for w in self.shards:
            t, n = C.c_int64(0), C.c_int64(0)
            # per-shard totals: read without re-running the monitor
            w._call("gbp_world_read_collision_totals", C.byref(t), C.byref(n))
            total += int(t.value)
            now += int(n.value)
        return total, now

    def read_robot_collisions(self):
        return np.concatenate([w.read_robot_collisions() for w in self.shards])

    def update_prior_of_horizon_state(self):
        self.shards[0].update_prior_of_horizon_state()

    def update_prior_of_current_state(self):
        self.shards[0].update_prior_of_current_state()

    def change_prior_of_variable(self, variable_index, robots, new_means):
        robots = np.asarray(robots, np.int64)
        new_means = np.asarray(new_means, np.float64).reshape(-1, 4)
        for q, w in enumerate(self.shards):
            lo, hi = self.bounds[q], self.bounds[q + 1]
            m = (robots >= lo) & (robots < hi)
            w.change_prior_of_variable(variable_index, (robots[m] - lo).astype(np.int32), new_means[m])

    def _by_shard(self, robots):
        robots = np.asarray(robots, np.int64)
        for q, w in enumerate(self.shards):
            lo, hi = self.bounds[q], self.bounds[q + 1]
            yield w, (robots >= lo) & (robots < hi), lo

    def reset_variables(self, robots, means, first_last_sigma=1e30, inbetween_sigma=float("inf")):
        robots = np.asarray(robots, np.int64)
        means = np.asarray(means, np.float64).reshape(robots.size, self.V, 4)
        for w, m, lo in self._by_shard(robots):
            w.reset_variables((robots[m] - lo).astype(np.int32), means[m], first_last_sigma, inbetween_sigma)

    def set_tracking_path(self, robots, paths):
        robots = np.asarray(robots, np.int64)
        for w, m, lo in self._by_shard(robots):
            if m.any():
                w.set_tracking_path((robots[m] - lo).astype(np.int32), [p for p, k in zip(paths, m) if k])

    def reset_tracking_factors(self, robots):
        robots = np.asarray(robots, np.int64)
        for w, m, lo in self._by_shard(robots):
            if m.any():
                w.reset_tracking_factors((robots[m] - lo).astype(np.int32))

    def remove_robots(self, robots):
        robots = np.asarray(robots, np.int64)
        for q, w in enumerate(self.shards):
            lo, hi = self.bounds[q], self.bounds[q + 1]
            m = (robots >= lo) & (robots < hi)
            w.remove_robots((robots[m] - lo).astype(np.int32))

    def read_removed(self):
        return np.concatenate([w.read_removed() for w in self.shards])

    def read_tracking(self):
        parts = [w.read_tracking() for w in self.shards]
        return tuple(np.concatenate([p[k] for p in parts], axis=0) for k in range(3))

    def iterate(self):
        self.shards[0].iterate()

    def iterate_schedule(self, internal, external):
        self.shards[0].iterate_schedule(internal, external)

    def internal_factor_iteration(self):
        self.shards[0].internal_factor_iteration()

    def internal_variable_iteration(self):
        self.shards[0].internal_variable_iteration()

    def external_factor_iteration(self):
        self.shards[0].external_factor_iteration()

    def external_variable_iteration(self):
        self.shards[0].external_variable_iteration()

    def step(self):
        self.shards[0].step()

    def change_factor_enabled(self, kind, enabled):
        for w in self.shards:
            w.change_factor_enabled(kind, enabled)

    def set_safety_distance_multiplier(self, m):
        for w in self.shards:
            w.set_safety_distance_multiplier(m)

    def set_schedule(self, kind, internal, external):
        for w in self.shards:
            w.set_schedule(kind, internal, external)

    # ---- read-back over the union, in global id order ---------------------------------------
    def read_beliefs(self, **kw):
        parts = [w.read_beliefs(**kw) for w in self.shards]
        return {k: np.concatenate([p[k] for p in parts], axis=0) for k in parts[0]}

    def read_positions(self):
        return np.concatenate([w.read_positions() for w in self.shards], axis=0)

    def read_connections(self):
        offs, nbs, rns, base = [np.zeros(1, np.int64)], [], [], 0
        for w in self.shards:
            o, nb, rn = w.read_connections()
            offs.append(o[1:] + base)
            nbs.append(nb)
            rns.append(rn)
            base += int(o[-1])
        return np.concatenate(offs), np.concatenate(nbs), np.concatenate(rns)

    def node_counts(self):
        return sum(w.node_counts() for w in self.shards)

    def sync(self):
        self.shards[0].sync()
