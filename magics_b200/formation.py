"""Scenario front-end of the hot path: `formation.yaml` -> the robots a run spawns.

Host-side restatement, in f32 where the reference is f32, of
  * `Formation::as_positions`   crates/gbp_config/src/formation.rs:304-454 (line-segment and circle shapes, the
    `Equal` and `Random` placement strategies, `Identity` / `Cross` projection of the waypoints) with its helpers
    `WorldDimensions::point_to_world_position` (:517-524), `polar` (:459-461),
    `randomly_place_nonoverlapping_circles_along_line_segment` (:548-591) and
    `evenly_place_nonoverlapping_circles_along_line_segment` (:595-641);
  * the spawn clock of `FormationSpawner` (crates/magics/src/planner/spawner.rs:186-323): first spawn when the
    formation's `delay` has elapsed, then once per `repeat.every`, `repeat.times` spawns in all;
  * the part of `spawn_formation` (spawner.rs:415-600) that turns positions into the route handed to
    `RobotBundle::new`: start pose, one pose per waypoint, velocities pointing at the next one.
The random draws (`rng.gen_range(0.0..1.0)` of rand 0.8.5 / wyrand, robot radii) are inputs here: the caller passes a
numpy Generator, so the layout is "a" legal draw of the reference's placement, not "the" draw of its PRNG stream.
The outputs feed `World.add_robots` and the oracle alike (magics_b200/scenarios.py `ReferenceScenario`).
"""
from __future__ import annotations

from dataclasses import dataclass, field

import numpy as np

f32 = np.float32
CURRENT, HORIZON, VARIABLE = 0, 1, 2   # CheckIntersectionWith (formation.rs:186-206) as gbp_reached_when_t codes
ROBOT_RADIUS, METER = 0, 1             # IntersectionDistance (formation.rs:150-166)


@dataclass
class ShapeSpec:
    kind: str                      # "line-segment" | "circle"
    points: tuple = ()             # line segment: ((x0, y0), (x1, y1)) in relative coordinates [0, 1]
    radius: float = 0.0            # circle
    center: tuple = (0.5, 0.5)


@dataclass
class Formation:
    robots: int
    delay_s: float
    repeat_every_s: float | None
    repeat_times: int | None       # None with repeat_every_s set: infinite
    initial_shape: ShapeSpec
    placement: str                 # "equal" | "random"
    attempts: int
    waypoints: list = field(default_factory=list)   # [(ShapeSpec, "identity" | "cross")]
    reached_when: tuple = (HORIZON, 0, ROBOT_RADIUS, 0.0)    # (intersects_with, variable index, distance kind, metres)
    finished_when: tuple = (HORIZON, 0, ROBOT_RADIUS, 0.0)
    planning_strategy: str = "only-local"

    def spawn_times(self, until_s: float) -> list[float]:
        """FormationSpawner: `delay`, then every `repeat.every`, `repeat.times` spawns in total (one without repeat)."""
        out, k = [], 0
        while True:
            t = self.delay_s + (k * self.repeat_every_s if self.repeat_every_s is not None else 0.0)
            if t > until_s:
                break
            out.append(t)
            k += 1
            if self.repeat_every_s is None or (self.repeat_times is not None and k >= self.repeat_times):
                break
            if self.repeat_every_s == 0.0:
                break
        return out


def _duration(d) -> float:
    return float(d.get("secs", 0)) + float(d.get("nanos", 0)) * 1e-9


def _shape(node) -> ShapeSpec:
    kind = node["kind"]
    if kind == "line-segment":
        a, b = node["value"]
        return ShapeSpec("line-segment", points=((float(a["x"]), float(a["y"])), (float(b["x"]), float(b["y"]))))
    if kind == "circle":
        c = node.get("center", {"x": 0.5, "y": 0.5})
        return ShapeSpec("circle", radius=float(node["radius"]), center=(float(c["x"]), float(c["y"])))
    raise ValueError(f"shape {kind!r}: Formation::as_positions has no arm for it (formation.rs:451)")


def _reached(node) -> tuple:
    if node is None:
        return (HORIZON, 0, ROBOT_RADIUS, 0.0)
    dist = node.get("distance", "robot-radius")
    dk, metres = (ROBOT_RADIUS, 0.0) if dist == "robot-radius" else (METER, float(dist["value"]))
    iw = node["intersects-with"]
    if iw == "current":
        return (CURRENT, 0, dk, metres)
    if iw == "horizon":
        return (HORIZON, 0, dk, metres)
    return (VARIABLE, int(iw["value"]), dk, metres)


def parse_formation_yaml(text: str) -> list[Formation]:
    """`formation.yaml` (serde_yaml, externally tagged enums: `!line-segment [..]`, `!random {attempts: n}`,
    `!finite n`, `!meter m`, `!variable i`)."""
    import yaml

    class Loader(yaml.SafeLoader):
        pass

    def tagged(loader, suffix, node):
        if isinstance(node, yaml.MappingNode):
            return {"kind": suffix, **loader.construct_mapping(node, deep=True)}
        if isinstance(node, yaml.SequenceNode):
            return {"kind": suffix, "value": loader.construct_sequence(node, deep=True)}
        return {"kind": suffix, "value": loader.construct_scalar(node)}

    Loader.add_multi_constructor("!", tagged)
    doc = yaml.load(text, Loader=Loader)
    return [formation_from_dict(f) for f in doc["formations"]]


def formation_from_dict(f: dict) -> Formation:
    rep = f.get("repeat")
    every = times = None
    if rep:
        every = _duration(rep["every"])
        t = rep.get("times", "infinite")
        # RepeatTimes (formation.rs:111-118): `infinite`, `!infinite` (unit variant as a tag) or `!finite N`
        times = None if t == "infinite" or (isinstance(t, dict) and t.get("kind") == "infinite") else int(t["value"])
    ps = f["initial-position"]["placement-strategy"]
    placement, attempts = ("equal", 0) if ps == "equal" else ("random", int(ps["attempts"]))
    return Formation(
        robots=int(f["robots"]), delay_s=_duration(f.get("delay", {})), repeat_every_s=every, repeat_times=times,
        initial_shape=_shape(f["initial-position"]["shape"]), placement=placement, attempts=attempts,
        waypoints=[(_shape(w["shape"]), w["projection-strategy"]) for w in f["waypoints"]],
        reached_when=_reached(f.get("waypoint-reached-when-intersects")),
        finished_when=_reached(f.get("finished-when-intersects")),
        planning_strategy=f.get("planning-strategy", "only-local"))


# ---- f32 helpers (glam 0.25 Vec2) ---------------------------------------------------------------
def _world(p, world_w: float, world_h: float) -> np.ndarray:
    """WorldDimensions::point_to_world_position: ((p - 0.5) * dim) in f64, then `as f32` (formation.rs:517-524)."""
    return np.array([f32((float(p[0]) - 0.5) * world_w), f32((float(p[1]) - 0.5) * world_h)], f32)


def _lerp(a: np.ndarray, b: np.ndarray, s) -> np.ndarray:
    """glam Vec2::lerp: self + ((rhs - self) * s)."""
    return (a + ((b - a).astype(f32) * f32(s)).astype(f32)).astype(f32)


def _length(v: np.ndarray) -> np.float32:
    """glam Vec2::length: sqrt(dot) with dot = x*x + y*y in f32."""
    return f32(np.sqrt(f32(f32(v[0] * v[0]) + f32(v[1] * v[1]))))


def _distance(a, b) -> np.float32:
    return _length((a - b).astype(f32))


def _randomly_along_segment(a, b, radii, attempts: int, rng) -> list | None:
    """randomly_place_nonoverlapping_circles_along_line_segment (formation.rs:548-591)."""
    n = len(radii)
    for _ in range(attempts):
        placed, lerps = [], []
        for radius in radii:
            amount = f32(rng.random(dtype=np.float32))  # rng.gen_range(0.0..1.0)
            pos = _lerp(a, b, amount)
            if all(_distance(pos, q) >= f32(r + f32(radius)) for q, r in placed):
                lerps.append(amount)
                placed.append((pos, f32(radius)))
                if len(placed) == n:
                    return lerps
    return None


def _evenly_along_segment(a, b, radii) -> list | None:
    """evenly_place_nonoverlapping_circles_along_line_segment (formation.rs:595-641), as written — the centre advances
    by `(r1 + diff) * 2 + (extra - diff) * dir`, a scalar added to both components of a vector."""
    rmin, rmax = f32(min(radii)), f32(max(radii))
    if f32(_distance(a, b) / rmax) < rmin:
        return None
    length = _distance(a, b)
    d = (b - a).astype(f32)
    dirn = (d * f32(f32(1.0) / _length(d))).astype(f32)  # glam normalize: self * length_recip()
    extra = f32(_distance(a, b) / rmax)
    center = (a + (f32(radii[0]) * dirn).astype(f32)).astype(f32)
    placed = []
    rs = [f32(r) for r in radii] + [f32(0.0)]
    for r1, r2 in zip(rs[:-1], rs[1:]):
        diff = f32(r2 - r1)
        placed.append(f32(_length((center - a).astype(f32)) / length))
        step = (f32(f32(r1 + diff) * f32(2.0)) + (f32(extra - diff) * dirn).astype(f32)).astype(f32)
        center = (center + step).astype(f32)
    return placed


def _polar(angle, magnitude) -> np.ndarray:
    """polar (formation.rs:459-461): f32 cos / sin of the angle times the magnitude."""
    return np.array([f32(np.cos(f32(angle))) * f32(magnitude), f32(np.sin(f32(angle))) * f32(magnitude)], f32)


def as_positions(fm: Formation, world_w: float, world_h: float, radii, rng):
    """Formation::as_positions -> (initial positions (n, 2) f32, [waypoint positions (n, 2) f32 per waypoint]) or None
    when the robots cannot be placed (formation.rs:304-454)."""
    n = fm.robots
    assert len(radii) == n
    sh = fm.initial_shape
    if sh.kind == "line-segment":
        a, b = _world(sh.points[0], world_w, world_h), _world(sh.points[1], world_w, world_h)
        lerps = (_randomly_along_segment(a, b, radii, fm.attempts, rng) if fm.placement == "random"
                 else _evenly_along_segment(a, b, radii))
        if lerps is None:
            return None
        assert len(lerps) == n
        init = np.stack([_lerp(a, b, s) for s in lerps]).astype(f32)
        wps = []
        for shape, proj in fm.waypoints:
            if shape.kind != "line-segment":
                raise NotImplementedError("no time for the other combinations sadly :( (formation.rs:372)")
            wa, wb = _world(shape.points[0], world_w, world_h), _world(shape.points[1], world_w, world_h)
            order = lerps if proj == "identity" else lerps[::-1]
            wps.append(np.stack([_lerp(wa, wb, s) for s in order]).astype(f32))
        return init, wps
    if sh.kind == "circle":
        if fm.placement != "equal":
            raise NotImplementedError("todo!() in the reference (formation.rs:405)")
        center = _world(sh.center, world_w, world_h)
        step = f32(f32(2.0) * f32(np.pi) / f32(n))
        angles = [f32(f32(i) * step) for i in range(n)]
        init = np.stack([(center + _polar(t, sh.radius)).astype(f32) for t in angles]).astype(f32)
        wps = []
        for shape, proj in fm.waypoints:
            if shape.kind != "circle":
                raise NotImplementedError("no time for the other combinations sadly :( (formation.rs:424)")
            if proj != "cross":
                raise ValueError("does not make sense for a circle (formation.rs:428)")
            c = _world(shape.center, world_w, world_h)
            wps.append(np.stack([(c + _polar(f32(t + f32(np.pi)), shape.radius)).astype(f32) for t in angles]).astype(f32))
        return init, wps
    raise NotImplementedError("Shape::Polygon: todo!() in the reference (formation.rs:451)")


def routes(init: np.ndarray, wps: list) -> list[np.ndarray]:
    """The waypoint polyline of every robot as spawn_formation builds it (spawner.rs:467-551): the start position
    followed by its position on each waypoint shape.  (The velocities the reference attaches — target speed toward the
    next point, the last pose inheriting the one before — are recomputed from these positions by
    scenarios.initial_means.)"""
    n = init.shape[0]
    return [np.stack([init[i]] + [w[i] for w in wps]).astype(f32) for i in range(n)]
