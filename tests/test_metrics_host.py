"""SURVEY §8(f) next-3, the consumer side of the export: the reference's evaluation metrics (scripts/ldj.py,
scripts/distance-travelled.py, scripts/perpendicular-path-deviation.py, scripts/utils.py) restated in
magics_b200/metrics.py against golden outputs of the reference's own functions (tests/golden/metrics.json, made by
tests/golden/make_golden_metrics.py), the Mission / Route clocks (planner/robot.rs:331-490, 815-1012) the exporter
reads, and an exported run (oracle-driven here; the engine's export is compared with the oracle's in
tests/test_gpu_evaluation.py) fed through the UNMODIFIED reference scripts where /root/reference exists.  CPU only."""
import json
import os
import subprocess
import sys

import numpy as np
import pytest

from magics_b200 import metrics, scenarios
from magics_b200.environment import Collider
from magics_b200.export import export_from_totals, format_color, obstacles_data
from magics_b200.mission import MissionClock, secs_f64

HERE = os.path.dirname(os.path.abspath(__file__))
GOLDEN = json.load(open(os.path.join(HERE, "golden", "metrics.json")))["cases"]
REL = 1e-12  # float64 formulas evaluated in a different order than the reference's numpy calls


def close(a, b, rel=REL):
    a, b = np.asarray(a, np.float64), np.asarray(b, np.float64)
    return a.shape == b.shape and np.all(np.abs(a - b) <= rel * np.maximum(1.0, np.maximum(np.abs(a), np.abs(b))))


@pytest.mark.parametrize("k", range(len(GOLDEN)))
def test_metrics_equal_the_reference_scripts_outputs(k):
    c = GOLDEN[k]
    pos, wps = np.array(c["positions"]), np.array(c["waypoints"])
    assert close(metrics.ldj(c["velocities"], c["timestamps"]), c["ldj"], 1e-10)  # log of a cancelling sum
    assert close(metrics.distance_travelled(pos), c["distance_travelled"])
    lines = np.array([metrics.closest_projection_onto_lines(p, wps) for p in pos])
    assert close(lines, c["closest_lines"], 1e-9)
    segs = np.array([metrics.closest_projection_onto_segments(p, wps) for p in pos])
    assert close(segs, c["closest_segments"], 1e-9)
    assert close(metrics.perpendicular_path_deviation(pos, wps, "lines"), c["deviation_lines"], 1e-9)
    assert close(metrics.perpendicular_path_deviation(pos, wps, "segments"), c["deviation_segments"], 1e-9)


def test_metrics_refuse_what_the_reference_asserts_against():
    with pytest.raises(ValueError):
        metrics.ldj(np.zeros((4, 2)), [0.0, 0.1, 0.1, 0.2])  # ldj.py:23 timestamps must increase
    with pytest.raises(ValueError):
        metrics.ldj(np.zeros((4, 3)), [0.0, 0.1, 0.2, 0.3])  # ldj.py:21 (n, 2)
    with pytest.raises(ValueError):
        metrics.distance_travelled(np.zeros((0, 2)))         # distance-travelled.py:31
    assert metrics.distance_travelled([[1.0, 2.0]]) == 0.0
    s = metrics.summary([3.0, 1.0, 2.0])
    assert s == {"robots": 3, "mean": 2.0, "median": 2.0, "largest": 3.0, "smallest": 1.0, "variance": 1.0, "stdev": 1.0}


def test_mission_clock_follows_route_advance_and_next_route():
    clock = MissionClock()
    clock.spawn([[(0, 0), (5, 0), (9, 0)], [(1, 1), (2, 2)]], started_at=1.5)
    assert clock.next_waypoint_index() == [1, 1]  # Route::new skips the initial pose (robot.rs:383)
    clock.observe([True, False], 2_100_000_000)
    assert clock.next_waypoint_index() == [2, 1] and not clock.missions[0].completed
    d = clock.mission_data(0, 3_000_000_000)
    assert d["waypoints"] == [[0.0, 0.0], [9.0, 0.0]]  # taskpoints of Mission::local: first and last (robot.rs:848)
    assert d["started_at"] == 1.5 and d["finished_at"] == 3.0 and d["routes"][0]["finished_at"] == 3.0  # still open: now
    clock.observe([True, True], 4_200_000_000)
    m0, m1 = clock.missions
    assert m0.completed and m1.completed and clock.next_waypoint_index() == [3, 2]
    # Route::advance adds the route's start to the (already absolute) clock (robot.rs:436-438); next_route does not (:975)
    assert m0.route.finished_at == secs_f64(4_200_000_000) + 1.5 and m0.finished_at == secs_f64(4_200_000_000)
    clock.observe([True, True], 9_000_000_000)  # MissionState::Completed: nothing moves any more (robot.rs:1004)
    assert clock.next_waypoint_index() == [3, 2] and m0.finished_at == secs_f64(4_200_000_000)
    d = clock.mission_data(0, 9_000_000_000)
    assert d["finished_at"] == m0.finished_at and len(d["routes"]) == 1 and len(d["routes"][0]["waypoints"]) == 3
    with pytest.raises(ValueError):
        clock.spawn([[(0, 0)]], 0.0)
    with pytest.raises(ValueError):
        clock.observe([True], 0)
    assert secs_f64(12_345_678_901) == 12.0 + 345_678_901 / 1e9  # Duration::as_secs_f64


def test_export_colour_and_obstacle_shapes():
    assert format_color(0x1e, 0x66, 0xf5) == "#1e66f5" and format_color(5, 0xe4, 0xf2) == "# 5e4f2"  # {:2x} pads with spaces
    obs = obstacles_data([Collider("ball", (1.0, 2.0), 0.0, radius=3.0),
                          Collider("cuboid", (10.0, -4.0), 0.0, half_extents=(2.0, 1.0)),
                          Collider("cuboid", (0.0, 0.0), float(np.pi / 2), half_extents=(2.0, 1.0)),
                          Collider("triangle", (7.0, 7.0), 0.3, points=((0, 0), (1, 0), (0, 1))),
                          Collider("convex-polygon", (5.0, 5.0), 0.0, points=((0, 0), (2, 0), (2, 2), (0, 2)))])
    assert obs["0"] == {"type": "Circle", "center": [1.0, 2.0], "radius": 3.0}
    assert obs["1"]["vertices"] == [[8.0, -5.0], [12.0, -5.0], [12.0, -3.0], [8.0, -3.0]]
    assert np.allclose(obs["2"]["vertices"], [[-1, -2], [1, -2], [1, 2], [-1, 2]], atol=1e-6)  # bounding box of the turned box
    assert obs["3"]["vertices"] == [[0.0, 0.0], [1.0, 0.0], [0.0, 1.0]]  # a Triangle's own vertices (export.rs:507-510)
    assert obs["4"]["vertices"] == [[5.0, 5.0], [7.0, 5.0], [7.0, 7.0], [5.0, 7.0]]


def _run_junction_with_the_oracle():
    """12 robots through the junction (3 waypoints each) with the trackers at 100 ms and the mission clock fed from
    `reached_waypoint`, on the CPU oracle (test infrastructure)."""
    from oracle.oracle import OracleWorld

    sw = scenarios.junction_twoway(per_lane=1)
    o = OracleWorld(sw.cfg, threads=4)
    sw.add_to(o)
    o.set_tracking_buffers(capacity=512, sample_ns=100_000_000)
    clock = MissionClock()
    wps = [sw.wp_xy[sw.wp_offsets[r]:sw.wp_offsets[r + 1]] for r in range(sw.n)]
    clock.spawn(wps, started_at=0.0)
    dt_ns = int(round(sw.cfg.delta_t * 1e9))
    task, fin = (2, 4, 1, 6.0), (2, 99, 1, 3.0)
    ticks = 150
    for tick in range(1, ticks + 1):
        reached = o.reached_waypoint(task, fin)
        clock.observe(reached, tick * dt_ns)
        assert clock.next_waypoint_index() == o.read_waypoint_index().tolist(), f"tick {tick}"
        o.step()
        o.update_robot_collisions()
        o.track(dt_ns, secs_f64(tick * dt_ns))
    totals = {"collisions_robots": o.read_robot_collisions(), "next_waypoint": o.read_waypoint_index(),
              "removed": np.zeros(sw.n, bool), "collisions_environment": None, "tracks": o.read_tracks(),
              "messages": None}
    colors = [format_color(30 + 17 * r, 102, 245 - 9 * r) for r in range(sw.n)]
    data = export_from_totals(totals, sw.n, sw.cfg, scenario="Structured Junction Twoway", makespan=ticks * dt_ns * 1e-9,
                              radii=sw.radii, missions=clock, now_ns=ticks * dt_ns, colors=colors,
                              colliders=[Collider("cuboid", (-30.0, -30.0), 0.0, half_extents=(20.0, 20.0))])
    return sw, clock, data, ticks * dt_ns


@pytest.fixture(scope="module")
def junction_export():
    return _run_junction_with_the_oracle()


def test_exported_run_evaluates(junction_export):
    sw, clock, data, now_ns = junction_export
    back = json.loads(json.dumps(data))
    assert set(back) >= {"scenario", "makespan", "delta_t", "gbp", "robots", "prng_seed", "obstacles"}
    finished = [m.completed for m in clock.missions]
    assert sum(finished) >= sw.n // 2, "most robots cross the junction in 15 s"
    for r, rd in enumerate(back["robots"].values()):
        assert set(rd) >= {"radius", "positions", "velocities", "collisions", "messages", "mission", "planning_strategy",
                           "color"}                                                      # RobotData (export.rs:112-123)
        ms = rd["mission"]
        assert set(ms) >= {"waypoints", "started_at", "finished_at", "routes"}           # MissionData (export.rs:125-132)
        assert len(ms["waypoints"]) == 2 and len(ms["routes"]) == 1 and len(ms["routes"][0]["waypoints"]) == 3
        assert ms["started_at"] == 0.0 and 0.0 < ms["finished_at"] <= secs_f64(now_ns)
        assert (ms["finished_at"] < secs_f64(now_ns)) == finished[r]
        assert len(rd["positions"]) == 150 and len(rd["velocities"]) == 149  # first velocity sample needs a previous position
    ev = metrics.evaluate(back)
    assert ev["makespan"] == pytest.approx(15.0) and ev["ldj"]["robots"] == sw.n
    for rid, e in ev["robots"].items():
        assert np.isfinite(e["ldj"]) and e["ldj"] < 0.0
        # start -> centre -> exit of a 90 m wide junction: between the straight line and a generous detour
        assert 30.0 < e["distance_travelled"] < 140.0, (rid, e)
    seg = metrics.evaluate(back, projection="segments")
    assert all(np.isfinite(e["path_deviation"]) and e["path_deviation"] < 3.0 for e in seg["robots"].values()), \
        "robots stay within a few metres of their waypoint polyline"


@pytest.mark.skipif(not os.path.isdir("/root/reference/scripts"), reason="the reference tree exists only in the build container")
def test_reference_scripts_read_the_export_unchanged(junction_export, tmp_path):
    """The reference's own consumers — scripts/ldj.py, scripts/distance-travelled.py and
    scripts/perpendicular-path-deviation.py, run as programs on the exported JSON — parse it and print, per robot, the
    numbers magics_b200.metrics computes from the same dict."""
    sw, clock, data, now_ns = junction_export
    path = tmp_path / "export_structured junction twoway_0.json"
    path.write_text(json.dumps(data))
    runner = (
        "import sys, runpy\n"
        "from unittest import mock\n"
        "for name in ('matplotlib', 'matplotlib.pyplot', 'toolz', 'toolz.curried', 'seaborn', 'result'):\n"
        "    sys.modules[name] = mock.MagicMock(name=name)\n"  # plotting only; not in this image
        "script = sys.argv[1]; sys.argv = sys.argv[1:]\n"
        "sys.path.insert(0, '/root/reference/scripts')\n"
        "runpy.run_path(script, run_name='__main__')\n")
    ev = metrics.evaluate(json.loads(path.read_text()))

    def run(script, *args):
        p = subprocess.run([sys.executable, "-c", runner, f"/root/reference/scripts/{script}", *args],
                           capture_output=True, text=True, timeout=120, env={**os.environ, "COLUMNS": "200"})
        assert p.returncode == 0, p.stderr[-2000:]
        return p.stdout

    def table_column(out, ids):
        """The per-robot table: rows `│ index │ id │ value │` in the order of the dict."""
        vals = {}
        for line in out.splitlines():
            cells = [c.strip() for c in line.replace("│", "|").split("|") if c.strip()]
            if len(cells) == 3 and cells[1] in ids and cells[0].isdigit():
                vals.setdefault(cells[1], float(cells[2]))
        return vals

    ids = set(data["robots"])
    got = table_column(run("ldj.py", "-i", str(path)), ids)
    assert set(got) == ids
    for rid in ids:
        assert abs(got[rid] - ev["robots"][rid]["ldj"]) <= 6e-4, (rid, got[rid], ev["robots"][rid]["ldj"])  # printed %.3f
    got = table_column(run("distance-travelled.py", str(path)), ids)
    assert set(got) == ids
    for rid in ids:
        assert abs(got[rid] - ev["robots"][rid]["distance_travelled"]) <= 6e-4
    got = table_column(run("perpendicular-path-deviation.py", "-i", str(path)), ids)
    assert set(got) == ids
    for rid in ids:
        want = ev["robots"][rid]["path_deviation"]
        assert (np.isnan(want) and np.isnan(got[rid])) or abs(got[rid] - want) <= 6e-4, (rid, got[rid], want)


def test_export_data_over_a_world_like_object_keeps_the_old_keys():
    """`export_data(world, ...)` as tests/test_gpu_evaluation.py calls it (no mission clock, no colliders): the keys of
    the first version of the export stay where they were."""
    from magics_b200.export import export_data

    class FakeWorld:
        num_robots = 2
        cfg = scenarios.circle(2, 8.0).cfg

        def export_totals(self):
            tr = (np.array([[0.0, 1.0], [2.0, 3.0]], np.float32), np.array([[4.0, 5.0]], np.float32), np.array([0.25]),
                  np.array([0.05]))
            return {"collisions_robots": np.array([1, 0], np.uint32), "next_waypoint": np.array([1, 2], np.int32),
                    "removed": np.array([False, True]), "collisions_environment": np.array([0, 3], np.uint32),
                    "tracks": [tr, tr], "messages": {"sent": {"internal": np.array([7, 8]), "external": np.array([1, 2])},
                                                      "received": {"internal": np.array([5, 6]), "external": np.array([3, 4])}}}

    d = json.loads(json.dumps(export_data(FakeWorld(), scenario="s", makespan=1.0, radii=[1.0, 2.0], prng_seed=3,
                                          waypoints=[[(0, 0), (1, 1)], [(2, 2), (3, 3)]])))
    r1 = d["robots"]["1"]
    assert r1["mission"] == {"waypoints": [[2.0, 2.0], [3.0, 3.0]], "next_waypoint": 2, "despawned": True}
    assert r1["collisions"] == {"robots": 0, "environment": 3} and r1["messages"]["received"] == {"internal": 6, "external": 4}
    assert r1["velocities"] == [{"velocity": [4.0, 0.0, 5.0], "timestamp": 0.25, "measured_over": {"secs": 0, "nanos": 50_000_000}}]
    assert d["obstacles"] == {} and d["prng_seed"] == 3 and d["gbp"]["iterations"]["internal"] == FakeWorld.cfg.iterations_internal
