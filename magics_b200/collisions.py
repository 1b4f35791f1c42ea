"""Which pairs collided, and where: the entry lists of `RobotRobotCollisions` / `RobotEnvironmentCollisions`.

The reference keeps a `CollisionHistory` per pair — robot-robot keyed by (entity r, entity c) in query order,
robot-environment by (robot, obstacle mesh) — and pushes the Aabb intersection of every Hit into it
(crates/magics/src/planner/collisions.rs:117-138, :417-426, :463-470, :700-716).  `export` writes one entry per pair
that ever hit, `collisions: {robots: [{robot_a, robot_b, aabbs}], environment: [{robot, obstacle, aabbs}]}`
(export.rs:171-214, :552-555); the thesis notebooks count those entries.

The engine's monitors (`gbp_world_update_robot_collisions`, `gbp_world_update_environment_collisions`) keep the state
machines on the device and return counters: total hits, pairs colliding now, hits per robot.  A Hit is a rare event, so
the entries are derived on the host from the counters that moved: the robots whose per-robot count went up are the only
candidates, the pair predicate is evaluated for them alone — robot-robot in numpy f32 in the kernel's operation order
(`k_robot_collisions`, gbp_shard.cuh), robot-environment by `gbp_collider_hits_ball`, the device predicate's own source
compiled for the host (gbp_collide_host.cpp) — and the result has to reproduce the device's counters exactly, otherwise
`CollisionLog` raises instead of exporting a guess.  Works on one world handle (a shard reports its own robots only).
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from .environment import pack_colliders


def _hits_ball(lib, arr, k, verts, nverts, xz, radii) -> np.ndarray:
    """intersection_test of collider k against the balls (xz[m, 2], radii[m]) -> bool[m]."""
    m = int(xz.shape[0])
    out = np.zeros(max(m, 1), np.uint8)
    xz = np.ascontiguousarray(xz, np.float32)
    radii = np.ascontiguousarray(radii, np.float32)
    rc = lib.gbp_collider_hits_ball(C.byref(arr[k]), C.c_int32(nverts), verts.ctypes.data_as(C.POINTER(C.c_float)),
                                    C.c_int32(m), xz.ctypes.data_as(C.POINTER(C.c_float)),
                                    radii.ctypes.data_as(C.POINTER(C.c_float)), out.ctypes.data_as(C.POINTER(C.c_uint8)))
    if rc != 0:
        raise RuntimeError(f"gbp_collider_hits_ball failed: {rc}")
    return out[:m].astype(bool)


def collider_aabbs(colliders) -> np.ndarray:
    """`Collider::aabb()` (gbp_global_planner/src/lib.rs:94-98) of every collider: (n, 4) f32 = mins x, z, maxs x, z."""
    from .world import load_library

    lib = load_library()
    arr, verts, _ = pack_colliders(colliders)
    nverts = int(verts.shape[0]) if any(c.points for c in colliders) else 0
    out = np.zeros((len(colliders), 4), np.float32)
    for k in range(len(colliders)):
        box = np.zeros(4, np.float32)
        rc = lib.gbp_collider_aabb(C.byref(arr[k]), C.c_int32(nverts), verts.ctypes.data_as(C.POINTER(C.c_float)),
                                   box.ctypes.data_as(C.POINTER(C.c_float)))
        if rc != 0:
            raise RuntimeError(f"gbp_collider_aabb failed: {rc}")
        out[k] = box
    return out


def ball_aabbs(xz: np.ndarray, radii: np.ndarray) -> np.ndarray:
    """parry2d `Ball::aabb(&Isometry2::translation(x, z))`: centre -+ radius, f32."""
    xz, r = np.asarray(xz, np.float32), np.asarray(radii, np.float32)[:, None]
    return np.concatenate([xz - r, xz + r], axis=1)


def aabb_intersection(a: np.ndarray, b: np.ndarray) -> np.ndarray:
    """parry2d `Aabb::intersection`: sup of the mins, inf of the maxs (kept as computed even if they cross)."""
    return np.concatenate([np.maximum(a[..., :2], b[..., :2]), np.minimum(a[..., 2:], b[..., 2:])], axis=-1)


def balls_intersect(xz_a, r_a, xz_b, r_b) -> np.ndarray:
    """parry2d `BoundingSphere::intersects` as `k_robot_collisions` evaluates it: |c_b - c_a|^2 <= (r_a + r_b)^2 with
    every f32 operation rounded on its own."""
    xz_a, xz_b = np.asarray(xz_a, np.float32), np.asarray(xz_b, np.float32)
    dx, dz = xz_b[..., 0] - xz_a[..., 0], xz_b[..., 1] - xz_a[..., 1]
    d2 = dx * dx + dz * dz
    sr = np.asarray(r_a, np.float32) + np.asarray(r_b, np.float32)
    return d2 <= sr * sr


class CollisionLog:
    """Feeds on a world's collision monitors; call `update_robot_collisions` / `update_environment_collisions` INSTEAD
    of the world's methods of the same name (every monitor update has to pass through the log, it forwards the call and
    returns what the world returned).  `radii` as given to `add_robots`; `colliders` as given to
    `set_environment_colliders` (call `set_colliders` when the environment is replaced)."""

    _CHUNK = 2048

    def __init__(self, world, radii, colliders=None):
        self.world = world
        self.radii = np.array(radii, np.float32).reshape(-1)
        self._rr_total = 0
        self._rr_per = np.zeros(0, np.int64)
        self._rr_colliding: set = set()
        self.robot_entries: dict = {}  # (r, c), r < c  ->  [aabb (4,) f32, ...] in hit order
        self.set_colliders(colliders)

    # ---- bookkeeping ---------------------------------------------------------------------------------------------
    def add_robots(self, radii):
        """Robots spawned later (their radii in spawn order)."""
        self.radii = np.concatenate([self.radii, np.array(radii, np.float32).reshape(-1)])

    def set_colliders(self, colliders):
        """A new `Colliders` resource: RobotEnvironmentCollisions::clear (collisions.rs:40-46)."""
        self.colliders = list(colliders or [])
        self._env_total = 0
        self._env_per = np.zeros(0, np.int64)
        self._env_colliding: set = set()
        self.environment_entries: dict = {}  # (robot, collider index) -> [aabb, ...]
        self._packed = None

    def _env_tools(self):
        if self._packed is None:
            from .world import load_library

            arr, verts, _ = pack_colliders(self.colliders)
            nverts = int(verts.shape[0]) if any(c.points for c in self.colliders) else 0
            self._packed = (load_library(), arr, verts, nverts, collider_aabbs(self.colliders))
        return self._packed

    @staticmethod
    def _grown(per: np.ndarray, n: int) -> np.ndarray:
        return per if per.shape[0] >= n else np.concatenate([per, np.zeros(n - per.shape[0], np.int64)])

    def _radii_for(self, n: int) -> np.ndarray:
        if self.radii.shape[0] < n:
            raise RuntimeError(f"CollisionLog knows {self.radii.shape[0]} radii, the world has {n} robots: add_robots()")
        return self.radii

    # ---- robot-robot ---------------------------------------------------------------------------------------------
    def update_robot_collisions(self):
        total, now = self.world.update_robot_collisions()
        if total == self._rr_total and not self._rr_colliding:
            if now != 0:
                raise RuntimeError(f"collision log out of step: {now} pairs colliding, none known")
            return total, now
        pos = self.world.read_positions()
        gone = np.asarray(self.world.read_removed()).astype(bool)
        rad = self._radii_for(pos.shape[0])
        # pairs known to be colliding: a despawned robot is no longer in the query, the others are re-tested
        still = set()
        for i, j in self._rr_colliding:
            if not (gone[i] or gone[j]) and bool(balls_intersect(pos[i], rad[i], pos[j], rad[j])):
                still.add((i, j))
        if total != self._rr_total:
            per = np.asarray(self.world.read_robot_collisions()).astype(np.int64)
            delta = per - self._grown(self._rr_per, per.shape[0])
            cand = np.nonzero(delta > 0)[0]
            cand = cand[~gone[cand]]
            new = []
            for lo in range(0, cand.shape[0], self._CHUNK):  # (chunk x all candidates) blocks of the pair predicate
                a = cand[lo:lo + self._CHUNK]
                hit = balls_intersect(pos[a][:, None, :], rad[a][:, None], pos[cand][None, :, :], rad[cand][None, :])
                ia, ib = np.nonzero(hit & (a[:, None] < cand[None, :]))
                new.extend((int(a[x]), int(cand[y])) for x, y in zip(ia, ib))
            new = sorted(p for p in new if p not in self._rr_colliding)  # was Free: CollisionStatus::Hit
            got = np.zeros_like(per)
            for i, j in new:
                got[i] += 1
                got[j] += 1
            if len(new) != total - self._rr_total or not np.array_equal(got, np.maximum(delta, 0)):
                raise RuntimeError(f"collision log out of step: {total - self._rr_total} new robot-robot hits reported, "
                                   f"{len(new)} pairs found")
            boxes = ball_aabbs(pos, rad)
            for i, j in new:  # r_aabb.intersection(&c_aabb), collisions.rs:121-125
                self.robot_entries.setdefault((i, j), []).append(aabb_intersection(boxes[i], boxes[j]))
                still.add((i, j))
            self._rr_total, self._rr_per = total, per
        if len(still) != now:
            raise RuntimeError(f"collision log out of step: {now} pairs colliding, {len(still)} known")
        self._rr_colliding = still
        return total, now

    # ---- robot-environment ---------------------------------------------------------------------------------------
    def update_environment_collisions(self):
        total, now = self.world.update_environment_collisions()
        if total == self._env_total and not self._env_colliding:
            if now != 0:
                raise RuntimeError(f"collision log out of step: {now} (robot, obstacle) pairs colliding, none known")
            return total, now
        lib, arr, verts, nverts, col_boxes = self._env_tools()
        pos = self.world.read_positions()
        gone = np.asarray(self.world.read_removed()).astype(bool)
        rad = self._radii_for(pos.shape[0])
        still = set()
        by_collider: dict = {}
        for r, c in self._env_colliding:
            if not gone[r]:
                by_collider.setdefault(c, []).append(r)
        for c, robots in by_collider.items():
            robots = np.asarray(sorted(robots))
            hit = _hits_ball(lib, arr, c, verts, nverts, pos[robots], rad[robots])
            still.update((int(r), c) for r in robots[hit])
        if total != self._env_total:
            per = np.asarray(self.world.read_environment_collisions()).astype(np.int64)
            delta = per - self._grown(self._env_per, per.shape[0])
            cand = np.nonzero(delta > 0)[0]
            cand = cand[~gone[cand]]
            new = []
            for c in range(len(self.colliders)):
                hit = _hits_ball(lib, arr, c, verts, nverts, pos[cand], rad[cand])
                new.extend((int(r), c) for r in cand[hit] if (int(r), c) not in self._env_colliding)
            new.sort()
            got = np.zeros_like(per)
            for r, _ in new:
                got[r] += 1
            if len(new) != total - self._env_total or not np.array_equal(got, np.maximum(delta, 0)):
                raise RuntimeError(f"collision log out of step: {total - self._env_total} new robot-environment hits "
                                   f"reported, {len(new)} pairs found")
            boxes = ball_aabbs(pos, rad)
            for r, c in new:  # robot_aabb.intersection(&env_aabb), collisions.rs:417-420
                self.environment_entries.setdefault((r, c), []).append(aabb_intersection(boxes[r], col_boxes[c]))
                still.add((r, c))
            self._env_total, self._env_per = total, per
        if len(still) != now:
            raise RuntimeError(f"collision log out of step: {now} (robot, obstacle) pairs colliding, {len(still)} known")
        self._env_colliding = still
        return total, now

    # ---- export ---------------------------------------------------------------------------------------------------
    @staticmethod
    def _aabb_json(box) -> dict:
        """serde of parry2d `Aabb { mins: Point2, maxs: Point2 }`."""
        return {"mins": [float(box[0]), float(box[1])], "maxs": [float(box[2]), float(box[3])]}

    def collision_data(self, robot_ids=None, obstacle_ids=None) -> dict:
        """`ExportData::collisions` (export.rs:208-214, :552-555): one entry per pair that ever hit, its Aabbs in hit
        order.  The reference iterates a HashMap (arbitrary order); entries come sorted by key here.  Ids default to
        robot index / position in the collider list, like the `robots` and `obstacles` tables of the export."""
        rid = (lambda r: r) if robot_ids is None else (lambda r: robot_ids[r])
        oid = (lambda c: c) if obstacle_ids is None else (lambda c: obstacle_ids[c])
        return {
            "robots": [{"robot_a": rid(i), "robot_b": rid(j), "aabbs": [self._aabb_json(b) for b in boxes]}
                       for (i, j), boxes in sorted(self.robot_entries.items())],
            "environment": [{"robot": rid(r), "obstacle": oid(c), "aabbs": [self._aabb_json(b) for b in boxes]}
                            for (r, c), boxes in sorted(self.environment_entries.items())],
        }
