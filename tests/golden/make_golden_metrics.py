"""Golden vectors for magics_b200/metrics.py, produced by the reference's own evaluation scripts.

Run in the build container (needs /root/reference; the fixture is committed so nothing reads the reference later):

  python tests/golden/make_golden_metrics.py        -> tests/golden/metrics.json

The reference's scripts are Python, so they are imported unmodified from /root/reference/scripts (by path: two of the
file names carry hyphens) and their functions are CALLED on seeded trajectories:
  ldj.py:18-56                               ldj(velocities, timestamps)
  distance-travelled.py:30-38                distance_travelled(positions)
  perpendicular-path-deviation.py:39-61      closest_projection_onto_line_segments(point, lines), the "rmse" of :117-118
  utils.py:150-196                           closest_projection_onto_line_segments(point, line points)
matplotlib / toolz / seaborn (plotting only, not in this image) are replaced by empty stand-ins before the import.
"""
import importlib.util
import json
import os
import sys
import types
from unittest import mock

import numpy as np

REF = "/root/reference/scripts"
HERE = os.path.dirname(os.path.abspath(__file__))


def load_reference_scripts():
    for name in ("matplotlib", "matplotlib.pyplot", "toolz", "toolz.curried", "seaborn", "result"):
        if name not in sys.modules:
            sys.modules[name] = mock.MagicMock(name=name)
    sys.path.insert(0, REF)  # utils.py does `from ldj import ldj`
    mods = {}
    for key, fname in (("ldj", "ldj.py"), ("distance", "distance-travelled.py"),
                       ("deviation", "perpendicular-path-deviation.py"), ("utils", "utils.py")):
        spec = importlib.util.spec_from_file_location("ref_" + key, os.path.join(REF, fname))
        m = importlib.util.module_from_spec(spec)
        with open(os.devnull, "w") as devnull, mock.patch("sys.stdout", devnull):  # two scripts print at import
            spec.loader.exec_module(m)
        mods[key] = m
    return types.SimpleNamespace(**mods)


def trajectories(rng):
    """Robot-like tracks: a route of 2-4 waypoints followed at varying speed with lateral wobble, sampled at the
    trackers' 100 ms period with a little jitter in the timestamps (velocity timestamps are f64 clock reads)."""
    out = []
    for k in range(8):
        nwp = 2 + k % 3
        wps = np.cumsum(rng.uniform(5.0, 40.0, size=(nwp, 2)) * rng.choice([-1.0, 1.0], size=(1, 2)), axis=0)
        wps += rng.uniform(-50, 50, size=(1, 2))
        n = int(rng.integers(12, 120))
        s = np.sort(rng.uniform(0, 1, size=n))
        s[0], s[-1] = 0.0, 1.0
        seglen = np.linalg.norm(np.diff(wps, axis=0), axis=1)
        cum = np.concatenate([[0.0], np.cumsum(seglen)]) / seglen.sum()
        pos = np.stack([np.interp(s, cum, wps[:, 0]), np.interp(s, cum, wps[:, 1])], axis=1)
        pos += rng.normal(0, 0.4, size=pos.shape)
        pos = pos.astype(np.float32).astype(np.float64)  # the exporter's positions are f32
        t = 0.1 * np.arange(1, n + 1) + rng.uniform(0, 1e-3, size=n)
        vel = np.gradient(pos, axis=0) / 0.1 + rng.normal(0, 0.05, size=pos.shape)
        vel = vel.astype(np.float32).astype(np.float64)
        out.append({"waypoints": wps.astype(np.float32).astype(np.float64).tolist(), "positions": pos.tolist(),
                    "velocities": vel.tolist(), "timestamps": t.tolist()})
    return out


def main():
    ref = load_reference_scripts()
    rng = np.random.default_rng(20240607)
    cases = trajectories(rng)
    for c in cases:
        wps = np.array(c["waypoints"])
        pos = np.array(c["positions"])
        c["ldj"] = float(ref.ldj.ldj(np.array(c["velocities"]), np.array(c["timestamps"])))
        c["distance_travelled"] = float(ref.distance.distance_travelled(pos))
        lines = [ref.deviation.line_from_line_segment(*s, *e) for s, e in zip(wps[:-1], wps[1:])]
        closest = np.array([ref.deviation.closest_projection_onto_line_segments(p, lines) for p in pos])
        c["closest_lines"] = closest.tolist()
        c["deviation_lines"] = float(np.sqrt(np.sum(np.linalg.norm(pos - closest, axis=1)) / len(pos)))
        lps = [ref.utils.LinePoints(start=list(s), end=list(e)) for s, e in zip(wps[:-1], wps[1:])]
        closest2 = np.array([ref.utils.closest_projection_onto_line_segments(p, lps) for p in pos])
        c["closest_segments"] = closest2.tolist()
        c["deviation_segments"] = float(np.sqrt(np.sum(np.linalg.norm(pos - closest2, axis=1)) / len(pos)))
    with open(os.path.join(HERE, "metrics.json"), "w") as f:
        json.dump({"source": "reference scripts/{ldj,distance-travelled,perpendicular-path-deviation,utils}.py, "
                             "called by tests/golden/make_golden_metrics.py", "cases": cases}, f)
    print(len(cases), "cases; ldj", [round(c["ldj"], 3) for c in cases])


if __name__ == "__main__":
    main()
