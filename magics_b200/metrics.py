"""The reference's evaluation metrics over an exported run ("makespan, collisions, LDJ").

The reference computes them in Python from the exporter's JSON: scripts/ldj.py:18-56 (log dimensionless jerk),
scripts/distance-travelled.py:30-38, scripts/perpendicular-path-deviation.py:39-61,104-118 (deviation from the
waypoint polyline; its "RMSE" is the root of the MEAN ABSOLUTE distance, kept as written), scripts/utils.py:150-196
(the projection the notebooks use, restricted to segments the point lies "between").  This module takes the same dict
(`magics_b200.export.export_data`) and returns the same numbers; tests/golden/metrics.json holds outputs of the
reference's own functions (tests/golden/make_golden_metrics.py imports them from /root/reference/scripts) on seeded
trajectories, and the tests compare to 1e-12 relative.
"""
from __future__ import annotations

import statistics

import numpy as np
from scipy.integrate import simpson


def _central_differences(y: np.ndarray, h: float) -> np.ndarray:
    """What `np.gradient(y, h)` returns for a uniform step: second-order central differences inside, first-order
    one-sided differences at both ends."""
    y = np.asarray(y, np.float64)
    if y.size < 2:
        raise ValueError("at least two samples")
    g = np.empty_like(y)
    g[1:-1] = (y[2:] - y[:-2]) / (2.0 * h)
    g[0] = (y[1] - y[0]) / h
    g[-1] = (y[-1] - y[-2]) / h
    return g


def ldj(velocities, timestamps) -> float:
    """Log dimensionless jerk of one robot (scripts/ldj.py:18-56).

    velocities (n, 2), timestamps (n,) strictly increasing.  The jerk is the second difference quotient of the velocity
    with the MEAN sample spacing; its squared norm is integrated by Simpson's rule over n equidistant abscissae spanning
    [t0, t1] (not the timestamps themselves), scaled by (t1 - t0)^3 / v_max^2."""
    v = np.asarray(velocities, np.float64)
    t = np.asarray(timestamps, np.float64)
    if v.ndim != 2 or v.shape[1] != 2 or len(v) == 0 or t.shape != (len(v),):
        raise ValueError("velocities (n, 2) and timestamps (n,) expected")
    if not np.all(np.diff(t) > 0) or not t[0] < t[-1]:
        raise ValueError("timestamps must increase strictly")
    h = float(np.mean(np.diff(t)))
    jerk_sq = np.zeros(len(v))
    for axis in (0, 1):
        j = _central_differences(_central_differences(v[:, axis], h), h)
        jerk_sq = jerk_sq + j * j if axis else j * j
    grid = np.linspace(t[0], t[-1], len(v))
    integral = simpson(jerk_sq, x=grid)
    v_max = np.max(np.sqrt(v[:, 0] ** 2 + v[:, 1] ** 2))
    return float(-np.log((t[-1] - t[0]) ** 3 / v_max ** 2 * integral))


def distance_travelled(positions) -> float:
    """Sum of the lengths of the sampled polyline (scripts/distance-travelled.py:30-38)."""
    p = np.asarray(positions, np.float64)
    if p.ndim != 2 or p.shape[1] != 2 or len(p) == 0:
        raise ValueError("positions (n, 2) expected")
    d = p[1:] - p[:-1]
    return float(np.sum(np.sqrt(np.sum(d * d, axis=1))))


def _route_waypoints(mission: dict) -> np.ndarray:
    """All waypoints of all routes in order (perpendicular-path-deviation.py:81-92)."""
    pts = [wp for route in mission["routes"] for wp in route["waypoints"]]
    return np.squeeze(np.asarray(pts, np.float64))


def closest_projection_onto_lines(point, waypoints) -> np.ndarray:
    """Nearest of the orthogonal projections of `point` onto the INFINITE lines y = a x + b through consecutive
    waypoints (perpendicular-path-deviation.py:39-61).  Vertical segments divide by zero there, as here."""
    x1, y1 = float(point[0]), float(point[1])
    best, best_d = None, None
    for (sx, sy), (ex, ey) in zip(waypoints[:-1], waypoints[1:]):
        with np.errstate(divide="ignore", invalid="ignore"):
            a = np.float64(ey - sy) / np.float64(ex - sx)
            b = sy - a * sx
            xp = (x1 + a * (y1 - b)) / (a * a + 1)
            proj = np.array([xp, a * xp + b])
        d = np.linalg.norm(proj - np.array([x1, y1]))
        if best is None or d < best_d:  # min() keeps the first of equals; NaN distances never win after a number
            best, best_d = proj, d
    return best


def _segment_is_valid(start, end, point) -> bool:
    """scripts/utils.py:162-169: the difference of the two bearings (atan2(dx, dy)) is at least a right angle."""
    v1, v2 = point - start, point - end
    return abs(np.arctan2(v1[0], v1[1]) - np.arctan2(v2[0], v2[1])) >= np.pi / 2


def closest_projection_onto_segments(point, waypoints) -> np.ndarray:
    """scripts/utils.py:171-196: projections onto the lines of the segments the point lies between (all of them if
    none qualifies, or if the nearest such projection is farther than 10 m)."""
    point = np.asarray(point, np.float64)
    segs = [(np.asarray(s, np.float64), np.asarray(e, np.float64)) for s, e in zip(waypoints[:-1], waypoints[1:])]

    def nearest(candidates):
        best, best_d = None, None
        for s, e in candidates:
            line = e - s
            proj = s + np.dot(point - s, line) / np.dot(line, line) * line
            d = np.linalg.norm(proj - point)
            if best is None or d < best_d:
                best, best_d = proj, d
        return best, best_d

    valid = [se for se in segs if _segment_is_valid(se[0], se[1], point)]
    best, best_d = nearest(valid if valid else segs)
    if best_d > 10:
        best, _ = nearest(segs)
    return best


def perpendicular_path_deviation(positions, waypoints, projection: str = "lines") -> float:
    """sqrt(sum of distances to the projection / number of samples) (perpendicular-path-deviation.py:117-118).
    projection: "lines" (that script) or "segments" (scripts/utils.py, used by the notebooks)."""
    p = np.asarray(positions, np.float64)
    w = np.asarray(waypoints, np.float64)
    project = {"lines": closest_projection_onto_lines, "segments": closest_projection_onto_segments}[projection]
    closest = np.array([project(q, w) for q in p])
    error = np.sum(np.linalg.norm(p - closest, axis=1))
    return float(np.sqrt(error / len(p)))


def summary(values) -> dict:
    """The table every script prints: robots, mean, median, largest, smallest, variance, stdev (`statistics`)."""
    v = [float(x) for x in values]
    out = {"robots": len(v), "mean": statistics.mean(v), "median": statistics.median(v), "largest": max(v),
           "smallest": min(v)}
    out["variance"] = statistics.variance(v) if len(v) > 1 else float("nan")
    out["stdev"] = statistics.stdev(v) if len(v) > 1 else float("nan")
    return out


def evaluate(data: dict, projection: str = "lines") -> dict:
    """Per-robot LDJ, distance travelled and path deviation of an export dict, with their summaries and the totals the
    notebooks read (makespan, collision counts)."""
    per_robot = {}
    for rid, rd in data["robots"].items():
        vel = np.asarray([m["velocity"] for m in rd["velocities"]], np.float64)
        ts = np.asarray([m["timestamp"] for m in rd["velocities"]], np.float64)
        pos = np.asarray(rd["positions"], np.float64)
        entry = {}
        if len(vel) >= 3:
            entry["ldj"] = ldj(vel[:, [0, 2]], ts)  # Bevy's (x, y-up, z): the plane is x-z (ldj.py:91)
        if len(pos) >= 1:
            entry["distance_travelled"] = distance_travelled(pos)
            if rd.get("mission", {}).get("routes"):
                entry["path_deviation"] = perpendicular_path_deviation(pos, _route_waypoints(rd["mission"]), projection)
        entry["collisions"] = dict(rd.get("collisions", {}))
        per_robot[rid] = entry
    out = {"robots": per_robot, "makespan": data.get("makespan")}
    for key in ("ldj", "distance_travelled", "path_deviation"):
        vals = [e[key] for e in per_robot.values() if key in e]
        if vals:
            out[key] = summary(vals)
    out["collisions"] = {k: sum(int(e["collisions"].get(k, 0)) for e in per_robot.values())
                         for k in ("robots", "environment")}
    if isinstance(data.get("collisions"), dict):
        # `collisions()` of the thesis notebooks (analyse-structured-junction-twoway.ipynb, analyse-collaborative-complex
        # .ipynb, analyse-environment-collisions.ipynb, analyse-comms-failure.ipynb): the NUMBER OF ENTRIES, i.e. of
        # pairs that ever hit — a pair that parts and hits again counts once here and twice in the totals above
        out["collision_entries"] = {"interrobot": len(data["collisions"].get("robots", [])),
                                    "environment": len(data["collisions"].get("environment", []))}
    return out
