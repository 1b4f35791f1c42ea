"""`ExportData` (crates/magics/src/export.rs:112-277, filled by `export` :277-470) assembled from the engine's read-backs.

The reference serialises, per robot: radius, positions (PositionTracker::positions), velocities
(VelocityTracker::measurements), collisions {robots, environment}, messages {sent, received} x {internal, external},
mission {waypoints, started_at, finished_at, routes}, planning_strategy, color; and globally scenario, makespan, delta_t,
gbp.iterations, prng_seed, config, obstacles, collisions, goal_areas.  What lives outside the iteration path (theme
colours, goal areas, the TOML config) is passed in by the caller or left out; the top-level `collisions` entry lists
(one entry per pair that ever hit, with the Aabb of every hit) come from a `magics_b200.collisions.CollisionLog`; the keys and
nesting of what is present follow the reference, and with a `magics_b200.mission.MissionClock` fed during the run the
reference's own consumers (scripts/ldj.py, scripts/distance-travelled.py, scripts/perpendicular-path-deviation.py) read
the JSON unchanged — tests/test_metrics_host.py runs them on it.  `magics_b200.metrics` computes the same numbers.
"""
from __future__ import annotations

import json

import numpy as np


def _duration(seconds: float) -> dict:
    """serde's `Duration` ({secs, nanos}) of `Duration::from_secs_f64` (nearest nanosecond)."""
    total = int(round(float(seconds) * 1e9))
    return {"secs": total // 1_000_000_000, "nanos": total % 1_000_000_000}


def format_color(r: int, g: int, b: int) -> str:
    """`format!("#{:2x}{:2x}{:2x}", r, g, b)` (export.rs:377): width 2 padded with SPACES, so a channel below 0x10
    comes out as e.g. "# 5e4f2" — kept, the plotting scripts get what the reference would hand them."""
    return "#{:2x}{:2x}{:2x}".format(int(r), int(g), int(b))


def obstacles_data(colliders) -> dict:
    """`ExportData::obstacles` (export.rs:503-548) from the `Colliders` handed to `set_environment_colliders`: a Ball
    becomes a Circle (centre = translation), a Triangle its own vertices, a ConvexPolygon its points shifted by the
    translation (the rotation is dropped there too), anything else the four corners of its bounding box.  Keyed by
    position in the list (the reference keys by the obstacle mesh's `Entity`)."""
    out = {}
    for k, c in enumerate(colliders or []):
        tx, ty = float(c.translation[0]), float(c.translation[1])
        if c.kind == "triangle":
            ob = {"type": "Polygon", "vertices": [[float(x), float(y)] for x, y in c.points]}
        elif c.kind == "ball":
            ob = {"type": "Circle", "center": [tx, ty], "radius": float(c.radius)}
        elif c.kind == "convex-polygon":
            ob = {"type": "Polygon", "vertices": [[float(x) + tx, float(y) + ty] for x, y in c.points]}
        else:  # cuboid: Aabb of the rotated box (parry2d Cuboid::aabb: |R| * half extents around the translation)
            ca, sa = abs(np.cos(np.float32(c.angle))), abs(np.sin(np.float32(c.angle)))
            hx = float(np.float32(ca * np.float32(c.half_extents[0]) + sa * np.float32(c.half_extents[1])))
            hy = float(np.float32(sa * np.float32(c.half_extents[0]) + ca * np.float32(c.half_extents[1])))
            ob = {"type": "Polygon", "vertices": [[tx - hx, ty - hy], [tx + hx, ty - hy], [tx + hx, ty + hy],
                                                  [tx - hx, ty + hy]]}
        out[str(k)] = ob
    return out


def export_from_totals(t: dict, n: int, cfg, *, scenario: str = "", makespan: float = 0.0, delta_t: float | None = None,
                       iterations: tuple | None = None, prng_seed: int = 0, radii=None, waypoints=None,
                       planning_strategy: str = "only-local", robot_ids=None, missions=None, now_ns: int | None = None,
                       colors=None, colliders=None, collision_log=None) -> dict:
    """Dict shaped like the reference's `ExportData` from the read-backs in `t` (`World.export_totals`).

    radii / waypoints: the per-robot inputs the caller gave `add_robots` (the engine does not read them back);
    robot_ids: the reference keys robots by Bevy `Entity`; any hashable ids, default 0..n-1;
    missions + now_ns: a `magics_b200.mission.MissionClock` fed during the run and the fixed clock at export time — fills
    `mission.started_at / finished_at / routes` (export.rs:381-409), which scripts/ldj.py and
    scripts/perpendicular-path-deviation.py read; colors: per-robot "#rrggbb" (`format_color`); colliders: see
    `obstacles_data`; collision_log: a `magics_b200.collisions.CollisionLog` fed during the run — fills the top-level
    `collisions: {robots, environment}` entry lists (export.rs:208-214, :552-555); robot and obstacle ids are the numbers whose
    strings key `robots` / `obstacles`."""
    ids = list(range(n)) if robot_ids is None else list(robot_ids)
    robots = {}
    for r in range(n):
        tracks = t["tracks"][r] if t.get("tracks") else (np.zeros((0, 2)), np.zeros((0, 2)), np.zeros(0), np.zeros(0))
        pos, vel, vt, vo = tracks
        msgs = t.get("messages")
        mission = {"waypoints": [[float(a), float(b)] for a, b in (waypoints[r] if waypoints is not None else [])]}
        if missions is not None:
            mission = missions.mission_data(r, 0 if now_ns is None else now_ns)
        mission["next_waypoint"] = int(t["next_waypoint"][r])
        mission["despawned"] = bool(t["removed"][r])
        robots[str(ids[r])] = {
            "radius": float(radii[r]) if radii is not None else None,
            "positions": [[float(a), float(b)] for a, b in pos],
            "velocities": [{"velocity": [float(v[0]), 0.0, float(v[1])], "timestamp": float(ts),
                            "measured_over": _duration(ov)} for v, ts, ov in zip(vel, vt, vo)],
            "collisions": {"robots": int(t["collisions_robots"][r]),
                           "environment": int(t["collisions_environment"][r])
                           if t.get("collisions_environment") is not None else 0},
            "messages": None if msgs is None else {
                "sent": {"internal": int(msgs["sent"]["internal"][r]), "external": int(msgs["sent"]["external"][r])},
                "received": {"internal": int(msgs["received"]["internal"][r]),
                             "external": int(msgs["received"]["external"][r])}},
            "mission": mission,
            "planning_strategy": planning_strategy,
            "color": colors[r] if colors is not None else "#000000",
        }
    it = iterations if iterations is not None else (cfg.iterations_internal, cfg.iterations_external)
    out = {"scenario": scenario, "makespan": float(makespan),
           "delta_t": float(cfg.delta_t if delta_t is None else delta_t),
           "gbp": {"iterations": {"internal": int(it[0]), "external": int(it[1])}},
           "robots": robots, "prng_seed": int(prng_seed), "obstacles": obstacles_data(colliders)}
    # `goal_areas`: the reference spawns none (the set-up system is commented out, goal_area.rs:8-11), so its exports
    # carry an empty map; kept for the shape
    out["goal_areas"] = {}
    if collision_log is not None:
        # `Entity` serialises as a number: scripts/plot-robot-positions.py:197-200 compares collision['obstacle'] with
        # int(key of `obstacles`), so the obstacle stays the integer position in the collider list
        out["collisions"] = collision_log.collision_data(robot_ids=None if robot_ids is None else ids)
    return out


def export_data(world, **kw) -> dict:
    """`export_from_totals` over a live world's read-backs; `json.dumps`-able."""
    return export_from_totals(world.export_totals(), world.num_robots, world.cfg, **kw)


def export_json(world, path: str, **kw) -> None:
    with open(path, "w") as f:
        json.dump(export_data(world, **kw), f)
