"""The entry lists of RobotRobotCollisions / RobotEnvironmentCollisions (planner/collisions.rs:117-138, :417-426,
:463-470, :700-716; exported by export.rs:171-214, :552-555) without a GPU.

`magics_b200.collisions.CollisionLog` derives them from a world's counters; here the world is the ORACLE (it has the
monitor methods of `magics_b200.World`), which also records every Hit with its Aabb itself — so the log's logic is
checked entry for entry, box for box.  The predicate the log asks about robot-environment pairs
(`gbp_collider_hits_ball`) and `gbp_collider_aabb` are host functions of the product library: they run here.
tests/test_gpu_zz_collision_pairs.py repeats the comparison with the engine as the world."""
import ctypes as C
import json

import numpy as np
import pytest

from magics_b200 import scenarios
from magics_b200.collisions import CollisionLog, collider_aabbs
from magics_b200.environment import Collider, Environment, pack_colliders, tile_colliders
from magics_b200.export import export_from_totals
from magics_b200.world import load_library
from oracle.oracle import OracleWorld


def entries_of(pairs, aabbs):
    out = {}
    for (a, b), box in zip(pairs.tolist(), aabbs):
        out.setdefault((a, b), []).append(box)
    return out


def assert_same_entries(log_entries, oracle_events, what):
    want = entries_of(*oracle_events)
    assert set(log_entries) == set(want), f"{what}: pairs {sorted(set(log_entries) ^ set(want))}"
    for k in want:
        assert len(log_entries[k]) == len(want[k]), (what, k)
        for got, ref in zip(log_entries[k], want[k]):
            assert np.array_equal(np.asarray(got, np.float32), ref), (what, k, got, ref)


SHAPES = [Collider("ball", (3.0, -2.0), 0.0, radius=1.25),
          Collider("cuboid", (-4.0, 1.0), 0.7, half_extents=(2.0, 0.5)),
          Collider("cuboid", (0.0, -6.0), 0.0, half_extents=(3.0, 1.0)),
          Collider("triangle", (0.5, 5.0), 0.7, points=((-1.0, -0.5), (2.0, -0.5), (0.3, 1.7))),
          Collider("convex-polygon", (6.0, 6.0), -0.7,
                   points=tuple((float(1.5 * np.cos(k * np.pi / 3)), float(1.5 * np.sin(k * np.pi / 3))) for k in range(6)))]


def test_host_predicate_of_the_library_equals_the_oracles_incl_rim_points():
    """gbp_collider_hits_ball (gbp_collide.cuh compiled for the host inside libgbp_b200.so) against the oracle's
    restatement of parry2d's intersection_test: random points and points at exactly one robot radius from a face."""
    rng = np.random.default_rng(3)
    R = np.float32(0.6)
    pts = rng.uniform(-9, 11, size=(5000, 2)).astype(np.float32)
    rim = [(3.0 + 1.25 + R, -2.0), (3.0, -2.0 + 1.25 + R), (3.0 + R, -6.0), (0.0, -5.0 + R), (-3.0 - R, -7.0 - R),
           (3.0 + R * np.float32(np.sqrt(0.5)), -5.0 + R * np.float32(np.sqrt(0.5)))]
    pts[: len(rim)] = np.asarray(rim, np.float32)
    sw = scenarios.circle(len(pts), 10.0, robot_radius=float(R))
    sw.positions[:] = pts
    o = OracleWorld(sw.cfg)
    sw.add_to(o)
    lib = load_library()
    arr, verts, _ = pack_colliders(SHAPES)
    radii = np.full(len(pts), R, np.float32)
    for k, c in enumerate(SHAPES):
        o.set_environment_colliders([c])
        o.update_environment_collisions()
        want = o.read_environment_collisions().astype(bool)
        out = np.zeros(len(pts), np.uint8)
        rc = lib.gbp_collider_hits_ball(C.byref(arr[k]), C.c_int32(verts.shape[0]), verts.ctypes.data_as(C.POINTER(C.c_float)),
                                        C.c_int32(len(pts)), pts.ctypes.data_as(C.POINTER(C.c_float)),
                                        radii.ctypes.data_as(C.POINTER(C.c_float)), out.ctypes.data_as(C.POINTER(C.c_uint8)))
        assert rc == 0
        assert np.array_equal(out.astype(bool), want), c.kind
        assert want.sum() > 20
    bad = arr[0].__class__(7, (C.c_float * 2)(0, 0), 0.0, 1.0, (C.c_float * 2)(0, 0), 0, 0)
    assert lib.gbp_collider_hits_ball(C.byref(bad), 0, None, 0, None, None, None) == -2  # GBP_ERR_BAD_ARGUMENT


def test_collider_aabbs_contain_their_shapes_tightly():
    """Collider::aabb against float64 geometry: ball centre -+ r, |R| half extents, min / max of the moved vertices."""
    boxes = collider_aabbs(SHAPES)
    for c, box in zip(SHAPES, boxes):
        a = float(np.float32(c.angle))
        rot = np.array([[np.cos(a), -np.sin(a)], [np.sin(a), np.cos(a)]])
        if c.kind == "ball":
            pts = np.asarray(c.translation) + c.radius * np.array([[-1, -1], [1, 1]])
        elif c.kind == "cuboid":
            hx, hy = c.half_extents
            pts = np.array([[-hx, -hy], [hx, -hy], [hx, hy], [-hx, hy]]) @ rot.T + np.asarray(c.translation)
        else:
            pts = np.asarray(c.points) @ rot.T + np.asarray(c.translation)
        want = np.concatenate([pts.min(axis=0), pts.max(axis=0)])
        assert np.allclose(box, want, rtol=0, atol=2e-6), (c.kind, box, want)


@pytest.mark.parametrize("interrobot", [0, 1])
def test_robot_robot_entries_equal_the_oracles(interrobot):
    """Circle swarm driving through its centre (InterRobot factors off: they collide, several pairs per tick, pairs
    that part and hit again), every tick through the log."""
    sw = scenarios.circle(8, circle_radius=8.0, robot_radius=1.0)
    sw.cfg.enable_interrobot = interrobot
    o = OracleWorld(sw.cfg)
    sw.add_to(o)
    log = CollisionLog(o, sw.radii)
    for tick in range(70):
        o.step()
        total, now = log.update_robot_collisions()
        if tick == 40:
            o.remove_robots([2])  # a despawned robot leaves the query; its pairs stop being updated
    assert_same_entries(log.robot_entries, o.read_collision_events(0), "robot-robot")
    if not interrobot:
        assert total >= 4 and len(log.robot_entries) >= 4
        assert sum(len(v) for v in log.robot_entries.values()) == total
    d = log.collision_data()
    assert [set(e) for e in d["robots"]] == [{"robot_a", "robot_b", "aabbs"}] * len(log.robot_entries)
    for e in d["robots"]:
        assert e["robot_a"] < e["robot_b"] and all(set(b) == {"mins", "maxs"} for b in e["aabbs"])
        assert all(b["mins"][0] <= b["maxs"][0] and b["mins"][1] <= b["maxs"][1] for b in e["aabbs"])


def test_robot_environment_entries_equal_the_oracles_with_robots_added_and_removed():
    """The '+' junction's tile colliders plus a ball in the middle, Obstacle factors off: robots drive over the walls,
    enter and leave colliders, a second swarm is spawned mid-run and two robots despawn."""
    env = Environment(grid=["┼"], tile_size=100.0, path_width=0.1325)
    cols = tile_colliders(env) + [Collider("ball", (0.0, 0.0), 0.0, radius=2.0)] + SHAPES[3:]
    sw = scenarios.circle(12, 45.0, robot_radius=1.0)
    o = OracleWorld(sw.cfg)
    sw.add_to(o)
    o.change_factor_enabled(2, 0)
    o.set_environment_colliders(cols)
    log = CollisionLog(o, sw.radii, cols)
    seen_now = []
    for tick in range(140):
        o.step()
        total, now = log.update_environment_collisions()
        log.update_robot_collisions()
        seen_now.append(now)
        if tick == 30:
            sw.add_to(o, set_sdf=False)
            log.add_robots(sw.radii)
        if tick == 60:
            o.remove_robots([1, 13])
    assert total >= 12 and max(seen_now) > min(seen_now)
    assert_same_entries(log.environment_entries, o.read_collision_events(1), "robot-environment")
    assert_same_entries(log.robot_entries, o.read_collision_events(0), "robot-robot")
    assert sum(len(v) for v in log.environment_entries.values()) == total
    # and in the export: top-level `collisions`, obstacles keyed like the `obstacles` table
    n = o.num_robots
    totals = {"collisions_robots": o.read_robot_collisions(), "next_waypoint": o.read_waypoint_index(),
              "removed": o.read_removed(), "collisions_environment": o.read_environment_collisions(), "tracks": None,
              "messages": None}
    d = json.loads(json.dumps(export_from_totals(totals, n, sw.cfg, scenario="junction", colliders=cols,
                                                 collision_log=log)))
    assert set(d["collisions"]) == {"robots", "environment"}
    assert len(d["collisions"]["environment"]) == len(log.environment_entries)
    assert all(str(e["obstacle"]) in d["obstacles"] and str(e["robot"]) in d["robots"] for e in d["collisions"]["environment"])
    # what the reference's consumers do with it: plot-robot-positions.py:197-200 (obstacle == int(key)), and the
    # notebooks' collisions() = number of entries
    hit_obstacles = {entity for entity in d["obstacles"] for c in d["collisions"]["environment"] if c["obstacle"] == int(entity)}
    assert hit_obstacles == {str(c) for _, c in log.environment_entries}
    from magics_b200 import metrics
    ev = metrics.evaluate(d)
    assert ev["collision_entries"] == {"interrobot": len(log.robot_entries), "environment": len(log.environment_entries)}
    assert ev["collisions"]["environment"] == total >= ev["collision_entries"]["environment"]
    # per-robot counts of the export are the sums over the robot's entries (RobotEnvironmentCollisions::get)
    for r in range(n):
        mine = sum(len(e["aabbs"]) for e in d["collisions"]["environment"] if e["robot"] == r)
        assert mine == d["robots"][str(r)]["collisions"]["environment"]


def test_a_monitor_update_that_bypasses_the_log_is_noticed():
    """Hits that began AND ended behind the log's back cannot be reconstructed: it raises instead of exporting a guess."""
    sw = scenarios.circle(8, circle_radius=8.0, robot_radius=1.0)
    sw.cfg.enable_interrobot = 0
    o = OracleWorld(sw.cfg)
    sw.add_to(o)
    log = CollisionLog(o, sw.radii)
    for tick in range(200):
        o.step()
        total, now = o.update_robot_collisions()  # not through the log
        if total > 0 and now == 0:
            break
    assert total > 0 and now == 0
    with pytest.raises(RuntimeError, match="out of step"):
        log.update_robot_collisions()


@pytest.mark.skipif(not __import__("os").path.isdir("/root/reference/scripts"),
                    reason="the reference tree exists only in the build container")
def test_reference_plot_script_reads_obstacles_and_collision_entries_unchanged(tmp_path):
    """scripts/plot-robot-positions.py, run as a program on an exported run with obstacles and collision entries: the
    one reference consumer that walks `collisions.environment` next to `obstacles` (:193-208, red = an obstacle that was
    hit).  matplotlib / plotly are not in this image and are mocked; the colour each obstacle patch gets is recorded."""
    import subprocess
    import sys

    from magics_b200.mission import MissionClock

    env = Environment(grid=["┼"], tile_size=100.0, path_width=0.1325)
    cols = tile_colliders(env) + [Collider("ball", (0.0, 0.0), 0.0, radius=2.0)]
    sw = scenarios.circle(6, 45.0, robot_radius=1.0)
    o = OracleWorld(sw.cfg)
    sw.add_to(o)
    o.change_factor_enabled(2, 0)
    o.set_environment_colliders(cols)
    o.set_tracking_buffers(capacity=256, sample_ns=100_000_000)
    clock = MissionClock()
    clock.spawn([sw.wp_xy[sw.wp_offsets[r]:sw.wp_offsets[r + 1]] for r in range(sw.n)], started_at=0.0)
    log = CollisionLog(o, sw.radii, cols)
    dt_ns = int(round(sw.cfg.delta_t * 1e9))
    for tick in range(1, 61):
        o.step()
        log.update_environment_collisions()
        log.update_robot_collisions()
        o.track(dt_ns, tick * dt_ns * 1e-9)
    assert log.environment_entries
    totals = {"collisions_robots": o.read_robot_collisions(), "next_waypoint": o.read_waypoint_index(),
              "removed": o.read_removed(), "collisions_environment": o.read_environment_collisions(),
              "tracks": o.read_tracks(), "messages": None}
    data = export_from_totals(totals, sw.n, sw.cfg, scenario="junction", radii=sw.radii, colliders=cols, missions=clock,
                              now_ns=60 * dt_ns, collision_log=log, colors=["#1e66f5"] * sw.n)
    path = tmp_path / "run.json"
    path.write_text(json.dumps(data))
    runner = (
        "import sys, runpy, json\n"
        "from unittest import mock\n"
        "mpl, plotly = mock.MagicMock(name='matplotlib'), mock.MagicMock(name='plotly')\n"
        "sys.modules.update({'matplotlib': mpl, 'matplotlib.pyplot': mpl.pyplot, 'plotly': plotly,\n"
        "                    'plotly.graph_objects': plotly.graph_objects})\n"
        "fig, ax = mock.MagicMock(), mock.MagicMock()\n"
        "mpl.pyplot.subplots.return_value = (fig, ax)\n"
        "script = sys.argv[1]; sys.argv = sys.argv[1:]\n"
        "runpy.run_path(script, run_name='__main__')\n"
        "patches = [c.kwargs.get('color') for c in mpl.pyplot.Circle.call_args_list + mpl.pyplot.Polygon.call_args_list]\n"
        "print('PATCH_COLOURS ' + json.dumps(patches))\n")
    p = subprocess.run([sys.executable, "-c", runner, "/root/reference/scripts/plot-robot-positions.py", "-i", str(path),
                        "-o", str(tmp_path / "out.svg")], capture_output=True, text=True, timeout=120)
    assert p.returncode == 0, p.stderr[-2000:]
    line = [ln for ln in p.stdout.splitlines() if ln.startswith("PATCH_COLOURS ")][0]
    colours = json.loads(line[len("PATCH_COLOURS "):])
    assert len(colours) == len(cols)
    assert colours.count("red") == len({c for _, c in log.environment_entries}) >= 1  # one red patch per obstacle hit


class _RandomWalkWorld:
    """A world made of nothing but the two monitors (numpy all-pairs state machines, robots on a random walk): dense,
    simultaneous events — robots in several pairs at once, many hits per update — which the oracle-driven runs above
    only touch lightly.  It keeps its own record of every Hit."""

    def __init__(self, n, colliders, rng, box=12.0):
        self.rng, self.box = rng, box
        self.pos = rng.uniform(-box, box, (n, 2)).astype(np.float32)
        self.radii = rng.uniform(0.4, 1.2, n).astype(np.float32)
        self.gone = np.zeros(n, bool)
        self.colliders = colliders
        self.rr_state, self.env_state = {}, {}
        self.rr_hits, self.env_hits = np.zeros(n, np.uint32), np.zeros(n, np.uint32)
        self.rr_total = self.env_total = 0
        self.rr_events, self.env_events = [], []
        self.lib = load_library()
        self.arr, self.verts, _ = pack_colliders(colliders)

    num_robots = property(lambda self: self.pos.shape[0])

    def move(self):
        self.pos = (self.pos + self.rng.normal(0.0, 0.6, self.pos.shape)).clip(-self.box, self.box).astype(np.float32)

    def read_positions(self):
        return self.pos.copy()

    def read_removed(self):
        return self.gone.astype(np.uint8)

    def read_robot_collisions(self):
        return self.rr_hits.copy()

    def read_environment_collisions(self):
        return self.env_hits.copy()

    def update_robot_collisions(self):
        from magics_b200.collisions import balls_intersect

        n, now_count = self.num_robots, 0
        for i in range(n):
            for j in range(i + 1, n):
                if self.gone[i] or self.gone[j]:
                    continue
                now = bool(balls_intersect(self.pos[i], self.radii[i], self.pos[j], self.radii[j]))
                if now and not self.rr_state.get((i, j), False):
                    self.rr_total += 1
                    self.rr_hits[i] += 1
                    self.rr_hits[j] += 1
                    self.rr_events.append((i, j))
                self.rr_state[(i, j)] = now
                now_count += now
        return self.rr_total, now_count

    def update_environment_collisions(self):
        n, now_count = self.num_robots, 0
        for c in range(len(self.colliders)):
            out = np.zeros(n, np.uint8)
            self.lib.gbp_collider_hits_ball(C.byref(self.arr[c]), C.c_int32(self.verts.shape[0]),
                                            self.verts.ctypes.data_as(C.POINTER(C.c_float)), C.c_int32(n),
                                            self.pos.ctypes.data_as(C.POINTER(C.c_float)),
                                            self.radii.ctypes.data_as(C.POINTER(C.c_float)),
                                            out.ctypes.data_as(C.POINTER(C.c_uint8)))
            for r in range(n):
                if self.gone[r]:
                    continue
                now = bool(out[r])
                if now and not self.env_state.get((r, c), False):
                    self.env_total += 1
                    self.env_hits[r] += 1
                    self.env_events.append((r, c))
                self.env_state[(r, c)] = now
                now_count += now
        return self.env_total, now_count


@pytest.mark.parametrize("seed", [0, 1, 2])
def test_dense_random_walk_every_hit_lands_in_the_right_entry(seed):
    rng = np.random.default_rng(seed)
    cols = SHAPES + [Collider("ball", (float(x), float(y)), 0.0, radius=1.0) for x, y in rng.uniform(-10, 10, (6, 2))]
    w = _RandomWalkWorld(40, cols, rng)
    log = CollisionLog(w, w.radii, cols)
    log._CHUNK = 7  # several blocks of the candidate x candidate pair test
    for tick in range(60):
        w.move()
        if tick == 30:
            w.gone[[3, 11]] = True
        assert log.update_robot_collisions()[0] == w.rr_total
        assert log.update_environment_collisions()[0] == w.env_total
    assert w.rr_total > 60 and w.env_total > 40  # dense: several hits per update
    want_rr, want_env = {}, {}
    for p in w.rr_events:
        want_rr[p] = want_rr.get(p, 0) + 1
    for p in w.env_events:
        want_env[p] = want_env.get(p, 0) + 1
    assert {k: len(v) for k, v in log.robot_entries.items()} == want_rr
    assert {k: len(v) for k, v in log.environment_entries.items()} == want_env
    assert max(want_rr.values()) >= 2  # pairs that parted and hit again
