"""`ExportData` (crates/magics/src/export.rs:112-277, filled by `export` :277-470) assembled from the engine's read-backs.

The reference serialises, per robot: radius, positions (PositionTracker::positions), velocities
(VelocityTracker::measurements), collisions {robots, environment}, messages {sent, received} x {internal, external},
mission {waypoints, started_at, finished_at, routes}, planning_strategy, color; and globally scenario, makespan, delta_t,
gbp.iterations, prng_seed, config, obstacles, collisions, goal_areas.  What lives outside the iteration path (theme
colours, goal areas, the TOML config, route timing kept by the mission system) is passed in by the caller or left out;
the keys and nesting of what is present follow the reference so that `scripts/ldj.py`-style consumers read it unchanged.
"""
from __future__ import annotations

import json

import numpy as np


def _duration(seconds: float) -> dict:
    """serde's `Duration` ({secs, nanos}) of `Duration::from_secs_f64` (nearest nanosecond)."""
    total = int(round(float(seconds) * 1e9))
    return {"secs": total // 1_000_000_000, "nanos": total % 1_000_000_000}


def export_data(world, *, scenario: str = "", makespan: float = 0.0, delta_t: float | None = None,
                iterations: tuple | None = None, prng_seed: int = 0, radii=None, waypoints=None,
                planning_strategy: str = "only-local", robot_ids=None) -> dict:
    """Dict shaped like the reference's `ExportData`; `json.dumps`-able.

    radii / waypoints: the per-robot inputs the caller gave `add_robots` (the engine does not read them back);
    robot_ids: the reference keys robots by Bevy `Entity`; any hashable ids, default 0..n-1."""
    t = world.export_totals()
    n = world.num_robots
    ids = list(range(n)) if robot_ids is None else list(robot_ids)
    robots = {}
    for r in range(n):
        tracks = t["tracks"][r] if t.get("tracks") else (np.zeros((0, 2)), np.zeros((0, 2)), np.zeros(0), np.zeros(0))
        pos, vel, vt, vo = tracks
        msgs = t.get("messages")
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
            "mission": {"waypoints": [[float(a), float(b)] for a, b in (waypoints[r] if waypoints is not None else [])],
                        "next_waypoint": int(t["next_waypoint"][r]), "despawned": bool(t["removed"][r])},
            "planning_strategy": planning_strategy,
        }
    cfg = world.cfg
    it = iterations if iterations is not None else (cfg.iterations_internal, cfg.iterations_external)
    return {"scenario": scenario, "makespan": float(makespan),
            "delta_t": float(cfg.delta_t if delta_t is None else delta_t),
            "gbp": {"iterations": {"internal": int(it[0]), "external": int(it[1])}},
            "robots": robots, "prng_seed": int(prng_seed)}


def export_json(world, path: str, **kw) -> None:
    with open(path, "w") as f:
        json.dump(export_data(world, **kw), f)
