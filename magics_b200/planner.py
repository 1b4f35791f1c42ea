"""A global planner for the `planning-strategy: rrt-star` formations of the headless runner (host side, CPU).

In the reference the search is `rrt::rrtstar::rrtstar` + `smooth_path` of the third-party `rrt` crate, wrapped by
gbp_global_planner/src/rrtstar.rs:15-89 and fed by `CollisionProblem` (gbp_global_planner/src/lib.rs:135-187).  The crate
is not in the reference tree and the search is random, so this file is NOT a restatement and carries no parity claim: it is
the published RRT* algorithm (Karaman & Frazzoli 2011: steer, choose the cheapest parent in the neighbourhood, rewire)
behind the reference's problem definition —

  is_feasible(p)   no collider of `Colliders` intersects a Ball(collision_radius) at p (parry2d intersection_test, here
                   `gbp_collider_hits_ball`, the predicate of the engine's environment-collision monitor);
  parameters       `[rrt]` of config.toml: step-size, neighbourhood-radius, collision-radius, max-iterations,
                   `[rrt.smoothing]` enabled / max-iterations / step-size;
  result           the path from start to end as a list of points, or None when the iteration budget runs out
                   (PathfindingError::ReachedMaxIterations -> the mission asks again, robot.rs:786-791).

The reference samples uniformly from [-2000, 2000]^2 whatever the world's size; here samples come from the bounding box of
the colliders (grown by two steps), with the goal itself every 20th sample — same tree, far fewer wasted draws.
`RRTStarPlanner(...)` is callable as `Simulation(global_planner=...)` expects: planner(start, end, colliders, rng).
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from .collisions import collider_aabbs
from .environment import pack_colliders


class Feasibility:
    """CollisionProblem::is_feasible over a fixed collider list (batched: points (m, 2) -> bool (m,))."""

    def __init__(self, colliders, collision_radius: float):
        from .world import load_library

        self.lib = load_library()
        self.colliders = list(colliders)
        self.radius = np.float32(collision_radius)
        self.arr, self.verts, _ = pack_colliders(self.colliders)
        self.nverts = int(self.verts.shape[0]) if any(c.points for c in self.colliders) else 0
        self.boxes = collider_aabbs(self.colliders) if self.colliders else np.zeros((0, 4), np.float32)

    def __call__(self, points) -> np.ndarray:
        pts = np.ascontiguousarray(np.asarray(points, np.float32).reshape(-1, 2))
        m = pts.shape[0]
        free = np.ones(m, bool)
        radii = np.full(m, self.radius, np.float32)
        out = np.zeros(max(m, 1), np.uint8)
        r = float(self.radius)
        for k in range(len(self.colliders)):
            b = self.boxes[k]  # cheap reject: the ball's box against the collider's Aabb
            near = (pts[:, 0] >= b[0] - r) & (pts[:, 0] <= b[2] + r) & (pts[:, 1] >= b[1] - r) & (pts[:, 1] <= b[3] + r) & free
            if not near.any():
                continue
            sub = np.ascontiguousarray(pts[near])
            rc = self.lib.gbp_collider_hits_ball(C.byref(self.arr[k]), C.c_int32(self.nverts),
                                                 self.verts.ctypes.data_as(C.POINTER(C.c_float)), C.c_int32(sub.shape[0]),
                                                 sub.ctypes.data_as(C.POINTER(C.c_float)),
                                                 radii.ctypes.data_as(C.POINTER(C.c_float)),
                                                 out.ctypes.data_as(C.POINTER(C.c_uint8)))
            if rc != 0:
                raise RuntimeError(f"gbp_collider_hits_ball failed: {rc}")
            idx = np.flatnonzero(near)
            free[idx[out[:sub.shape[0]].astype(bool)]] = False
        return free

    def segment(self, a, b, step: float) -> bool:
        """Every point of a -> b at `step` spacing (both ends included) is feasible."""
        a, b = np.asarray(a, np.float64), np.asarray(b, np.float64)
        n = max(1, int(np.ceil(np.linalg.norm(b - a) / step)))
        t = np.linspace(0.0, 1.0, n + 1)[:, None]
        return bool(self(a + (b - a) * t).all())


def rrt_star(start, end, feasible: Feasibility, rng, *, step_size: float, neighbourhood_radius: float,
             max_iterations: int, bounds, goal_bias: int = 20, check_step: float | None = None):
    """RRT* from start to end; stops when a node lands within one step of the goal and the last edge is free
    (`stop_when_reach_goal = true` in the reference's call).  Returns the list of points start ... end or None."""
    start, end = np.asarray(start, np.float64), np.asarray(end, np.float64)
    check_step = check_step or step_size / 4.0
    cap = 1024
    pts = np.zeros((cap, 2))
    parent = np.full(cap, -1, np.int64)
    cost = np.zeros(cap)
    pts[0] = start
    n = 1
    lo, hi = np.asarray(bounds[0], np.float64), np.asarray(bounds[1], np.float64)
    for it in range(int(max_iterations)):
        q = end if goal_bias and it % goal_bias == goal_bias - 1 else lo + (hi - lo) * rng.random(2)
        d = np.linalg.norm(pts[:n] - q, axis=1)
        near = int(np.argmin(d))
        if d[near] == 0.0:
            continue
        new = q if d[near] <= step_size else pts[near] + (q - pts[near]) * (step_size / d[near])
        if not feasible(new)[0]:
            continue
        dn = np.linalg.norm(pts[:n] - new, axis=1)
        hood = np.flatnonzero(dn <= neighbourhood_radius)
        best, best_cost = -1, np.inf
        for j in hood[np.argsort(cost[hood] + dn[hood])]:  # cheapest parent whose edge is free
            if feasible.segment(pts[j], new, check_step):
                best, best_cost = int(j), cost[j] + dn[j]
                break
        if best < 0:
            continue
        if n == cap:
            cap *= 2
            pts, parent, cost = np.resize(pts, (cap, 2)), np.resize(parent, cap), np.resize(cost, cap)
        pts[n], parent[n], cost[n] = new, best, best_cost
        for j in hood:  # rewire
            if j != best and best_cost + dn[j] < cost[j] and feasible.segment(new, pts[j], check_step):
                parent[j], cost[j] = n, best_cost + dn[j]
        n += 1
        if np.linalg.norm(new - end) <= step_size and feasible.segment(new, end, check_step):
            path, k = [end], n - 1
            while k >= 0:
                path.append(pts[k].copy())
                k = int(parent[k])
            return [tuple(map(float, p)) for p in path[::-1]]
    return None


def smooth_path(path, feasible: Feasibility, rng, *, step_size: float, max_iterations: int):
    """Random shortcutting (what `rrt::smooth_path` does): pick two points on the path, join them if the segment is free."""
    path = [np.asarray(p, np.float64) for p in path]
    for _ in range(int(max_iterations)):
        if len(path) <= 2:
            break
        i, j = sorted(rng.integers(0, len(path), 2))
        if j - i < 2:
            continue
        if feasible.segment(path[i], path[j], step_size):
            path = path[:i + 1] + path[j:]
    return [tuple(map(float, p)) for p in path]


class RRTStarPlanner:
    """planner(start, end, colliders, rng) for `Simulation(global_planner=...)`, parameters as in `[rrt]` of config.toml."""

    def __init__(self, step_size: float = 5.0, collision_radius: float = 3.0, neighbourhood_radius: float = 8.0,
                 max_iterations: int = 50_000, smoothing: bool = True, smoothing_iterations: int = 500,
                 smoothing_step_size: float = 0.5):
        self.step_size, self.collision_radius = float(step_size), float(collision_radius)
        self.neighbourhood_radius, self.max_iterations = float(neighbourhood_radius), int(max_iterations)
        self.smoothing, self.smoothing_iterations = bool(smoothing), int(smoothing_iterations)
        self.smoothing_step_size = float(smoothing_step_size)
        self._feasible = None
        self._for = None

    @classmethod
    def from_config(cls, rrt: dict, max_iterations_cap: int = 200_000) -> "RRTStarPlanner":
        """From the `[rrt]` table of a scenario's config.toml (the reference's 5 000 000 iterations are capped)."""
        sm = rrt.get("smoothing", {})
        return cls(step_size=rrt.get("step-size", 5.0), collision_radius=rrt.get("collision-radius", 3.0),
                   neighbourhood_radius=rrt.get("neighbourhood-radius", 8.0),
                   max_iterations=min(int(rrt.get("max-iterations", 50_000)), max_iterations_cap),
                   smoothing=sm.get("enabled", True), smoothing_iterations=sm.get("max-iterations", 500),
                   smoothing_step_size=sm.get("step-size", 0.5))

    def __call__(self, start, end, colliders, rng):
        if self._for is not colliders:
            self._feasible, self._for = Feasibility(colliders, self.collision_radius), colliders
        f = self._feasible
        pad = 2.0 * self.step_size
        corners = np.concatenate([f.boxes[:, :2], f.boxes[:, 2:], np.asarray([start, end], np.float32)])
        bounds = (corners.min(axis=0) - pad, corners.max(axis=0) + pad)
        path = rrt_star(start, end, f, rng, step_size=self.step_size, neighbourhood_radius=self.neighbourhood_radius,
                        max_iterations=self.max_iterations, bounds=bounds)
        if path is not None and self.smoothing:
            path = smooth_path(path, f, rng, step_size=self.smoothing_step_size, max_iterations=self.smoothing_iterations)
        return path
