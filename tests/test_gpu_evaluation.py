"""SURVEY §8 next-3 on the device: robot-environment collisions (planner/collisions.rs:368-455), the position /
velocity sample buffers (planner/tracking.rs:117-260) and the export shape (export.rs:112-277) — engine == oracle,
bit for bit, on the same inputs."""
import json

import numpy as np
import pytest

from magics_b200 import World, scenarios
from magics_b200.environment import Collider, Environment, tile_colliders
from magics_b200.export import export_data
from oracle.oracle import OracleWorld

pytestmark = pytest.mark.gpu


def _both(sw):
    g, o = World(sw.cfg), OracleWorld(sw.cfg)
    sw.add_to(g)
    sw.add_to(o)
    return g, o


def test_every_shape_kind_random_points_and_rim_points():
    rng = np.random.default_rng(11)
    tri = ((-1.0, -0.5), (2.0, -0.5), (0.3, 1.7))
    hexagon = tuple((float(1.5 * np.cos(k * np.pi / 3)), float(1.5 * np.sin(k * np.pi / 3))) for k in range(6))
    cols = [Collider("ball", (3.0, -2.0), 0.0, radius=1.25),
            Collider("cuboid", (-4.0, 1.0), 0.7, half_extents=(2.0, 0.5)),
            Collider("cuboid", (0.0, -6.0), 0.0, half_extents=(3.0, 1.0)),
            Collider("triangle", (0.5, 5.0), 0.7, points=tri),
            Collider("convex-polygon", (6.0, 6.0), -0.7, points=hexagon)]
    cols += [Collider("ball", (float(x), float(y)), 0.0, radius=0.3) for x, y in rng.uniform(-9, 11, (40, 2))]  # > 32
    pts = rng.uniform(-9, 11, size=(6000, 2)).astype(np.float32)
    # points at exactly robot radius from an axis-aligned face, a corner, a ball: the f32 rim
    R = np.float32(0.6)
    rim = [(3.0 + 1.25 + R, -2.0), (3.0, -2.0 + 1.25 + R), (3.0 + R, -6.0), (0.0, -5.0 + R), (-3.0 - R, -7.0 - R),
           (3.0 + R * np.float32(np.sqrt(0.5)), -5.0 + R * np.float32(np.sqrt(0.5)))]
    pts[: len(rim)] = np.asarray(rim, np.float32)
    sw = scenarios.circle(len(pts), 10.0, robot_radius=float(R))
    sw.positions[:] = pts
    g, o = _both(sw)
    for w in (g, o):
        w.set_environment_colliders(cols)
    assert g.update_environment_collisions() == o.update_environment_collisions()
    hg, ho = g.read_environment_collisions(), o.read_environment_collisions()
    assert np.array_equal(hg, ho) and hg.sum() > 300
    assert g.update_environment_collisions() == o.update_environment_collisions()  # nothing new: same totals
    assert np.array_equal(g.read_environment_collisions(), ho)
    # one collider at a time (per shape kind), so that a compensating pair of errors cannot hide
    for c in cols[:5]:
        for w in (g, o):
            w.set_environment_colliders([c])
            w.update_environment_collisions()
        assert np.array_equal(g.read_environment_collisions(), o.read_environment_collisions()), c.kind


def test_junction_walls_over_a_run_with_robots_added_and_removed():
    """The '+' junction's tile colliders (map_generator.rs:537-1298) and robots that are pushed through the walls:
    the Obstacle factors are switched off, so the swarm drives straight over them and collisions begin and end."""
    env = Environment(grid=["┼"], tile_size=100.0, path_width=0.1325)
    cols = tile_colliders(env) + [Collider("ball", (0.0, 0.0), 0.0, radius=2.0)]
    sw = scenarios.circle(12, 45.0, robot_radius=1.0)
    g, o = _both(sw)
    for w in (g, o):
        w.change_factor_enabled(2, 0)
        w.set_environment_colliders(cols)
    totals = []
    for tick in range(140):
        g.step()
        o.step()
        tg, to = g.update_environment_collisions(), o.update_environment_collisions()
        assert tg == to, f"tick {tick}"
        totals.append(tg)
        if tick == 30:
            for w in (g, o):
                sw.add_to(w, set_sdf=False)
        if tick == 60:
            for w in (g, o):
                w.remove_robots([1, 13])
        if tick % 20 == 19:
            assert np.array_equal(g.read_environment_collisions(), o.read_environment_collisions()), f"tick {tick}"
    assert np.array_equal(g.read_positions(), o.read_positions())
    assert totals[-1][0] >= 12 and max(t[1] for t in totals) > min(t[1] for t in totals)  # entered and left walls


def test_trackers_match_sample_for_sample():
    sw = scenarios.circle(9, 12.0)
    g, o = _both(sw)
    for w in (g, o):
        w.set_tracking_buffers(capacity=6, sample_ns=100_000_000)
    dt_ns = 33_333_333
    for tick in range(1, 61):
        g.step()
        o.step()
        if tick == 25:
            idle = np.zeros(sw.n, np.uint8)
            idle[[2, 5]] = 1
            for w in (g, o):
                w.set_comms(idle=idle)
        if tick == 40:
            for w in (g, o):
                sw.add_to(w, set_sdf=False)
        for w in (g, o):
            w.track(dt_ns, tick * (dt_ns * 1e-9))
    tg, to = g.read_tracks(), o.read_tracks()
    assert len(tg) == len(to) == 2 * sw.n
    for r, (a, b) in enumerate(zip(tg, to)):
        for x, y, what in zip(a, b, ("positions", "velocities", "timestamps", "measured_over")):
            assert x.shape == y.shape and np.array_equal(x, y), (r, what)
    # full ring; the idle robot's samples stop when it stops; the late robot has fewer than a full ring
    assert len(tg[0][0]) == 6 and tg[2][2][-1] < tg[0][2][-1] - 0.5 and 1 <= len(tg[sw.n][0]) <= 6


def test_export_data_has_the_reference_shape():
    sw = scenarios.circle(5, 8.0)
    g = World(sw.cfg)
    g.set_message_counting(True)  # before the first robot: the reference counts from the creation of the graph on
    sw.add_to(g)
    g.set_tracking_buffers(capacity=8, sample_ns=50_000_000)
    g.set_environment_colliders([Collider("ball", (0.0, 0.0), 0.0, radius=1.0)])
    for tick in range(1, 31):
        g.step()
        g.update_robot_collisions()
        g.update_environment_collisions()
        g.track(20_000_000, tick * 0.02)
    d = export_data(g, scenario="circle-5", makespan=0.6, radii=sw.radii, prng_seed=3)
    text = json.dumps(d)
    back = json.loads(text)
    assert set(back) >= {"scenario", "makespan", "delta_t", "gbp", "robots", "prng_seed"}
    r0 = back["robots"]["0"]
    assert set(r0) >= {"radius", "positions", "velocities", "collisions", "messages", "mission", "planning_strategy"}
    assert set(r0["collisions"]) == {"robots", "environment"}
    assert set(r0["messages"]["sent"]) == {"internal", "external"} and r0["messages"]["sent"]["internal"] > 0
    assert len(r0["positions"]) == 8 and set(r0["velocities"][0]) == {"velocity", "timestamp", "measured_over"}
    assert back["gbp"]["iterations"] == {"internal": sw.cfg.iterations_internal, "external": sw.cfg.iterations_external}


def test_exported_mission_run_equals_the_oracles_and_evaluates():
    """A whole mission run exported: 12 robots through the junction (3 waypoints each), `reached_waypoint` on the device
    feeding the host's Mission / Route clocks (magics_b200/mission.py; robot.rs:331-490, 815-1012), robot-robot monitor
    and 100 ms trackers every tick.  The engine's `ExportData` equals the one assembled from the oracle's read-backs key
    for key, and the reference's evaluation metrics (magics_b200/metrics.py; pinned to the reference's own scripts in
    tests/test_metrics_host.py) come out of it."""
    from magics_b200 import metrics
    from magics_b200.export import export_from_totals
    from magics_b200.mission import MissionClock, secs_f64

    sw = scenarios.junction_twoway(per_lane=1)
    g, o = _both(sw)
    for w in (g, o):
        w.set_tracking_buffers(capacity=256, sample_ns=100_000_000)
    clock = MissionClock()
    clock.spawn([sw.wp_xy[sw.wp_offsets[r]:sw.wp_offsets[r + 1]] for r in range(sw.n)], started_at=0.0)
    dt_ns = int(round(sw.cfg.delta_t * 1e9))
    task, fin = (2, 4, 1, 6.0), (2, 99, 1, 3.0)
    ticks = 150
    for tick in range(1, ticks + 1):
        rg, ro = g.reached_waypoint(task, fin), o.reached_waypoint(task, fin)
        assert np.array_equal(rg, ro), f"tick {tick}"
        clock.observe(rg, tick * dt_ns)
        for w in (g, o):
            w.step()
            w.update_robot_collisions()
            w.track(dt_ns, secs_f64(tick * dt_ns))
    assert clock.next_waypoint_index() == g.read_waypoint_index().tolist() == o.read_waypoint_index().tolist()
    kw = dict(scenario="Structured Junction Twoway", makespan=ticks * dt_ns * 1e-9, radii=sw.radii, missions=clock,
              now_ns=ticks * dt_ns)
    dg = export_data(g, **kw)
    oracle_totals = {"collisions_robots": o.read_robot_collisions(), "next_waypoint": o.read_waypoint_index(),
                     "removed": np.zeros(sw.n, bool), "collisions_environment": None, "tracks": o.read_tracks(),
                     "messages": None}
    do = export_from_totals(oracle_totals, sw.n, sw.cfg, **kw)
    assert json.loads(json.dumps(dg)) == json.loads(json.dumps(do))
    r0 = dg["robots"]["0"]
    assert len(r0["positions"]) == ticks and set(r0["mission"]) >= {"waypoints", "started_at", "finished_at", "routes"}
    ev = metrics.evaluate(dg, projection="segments")
    assert ev["ldj"]["robots"] == sw.n and all(np.isfinite(e["ldj"]) and 30.0 < e["distance_travelled"] < 140.0 and
                                               e["path_deviation"] < 3.0 for e in ev["robots"].values())
    assert sum(m.completed for m in clock.missions) >= sw.n // 2


@pytest.mark.parametrize("name, obstacle_factors", [("Collaborative Complex", True), ("Structured Junction Twoway", False)])
def test_reference_scenarios_with_their_tile_colliders_and_trackers(name, obstacle_factors):
    """The reference's own scenario inputs (tests/golden/scenarios.json) with the evaluation systems running every tick:
    the tile colliders of the scenario's environment (map_generator.rs:537-1298), the environment-collision monitor and
    the position / velocity trackers at the reference's 100 ms period.  The junction run also gets a bollard on the
    crossing (and no Obstacle factors), so that collisions begin and end; either way engine == oracle."""
    from oracle import oracle as O

    sc = scenarios.ReferenceScenario(name)
    g, o = World(sc.cfg), OracleWorld(sc.cfg)
    g.set_sdf_from_environment(sc.env)
    o.set_sdf(O.env_to_sdf_image(sc.env))
    cols = tile_colliders(sc.env)
    assert len(cols) > 3
    if not obstacle_factors:
        cols = cols + [Collider("ball", (0.0, 0.0), 0.0, radius=4.0)]  # a bollard on the crossing: the lanes pass it
    rng = np.random.default_rng(0)
    ticks = 60 if obstacle_factors else 95  # the first robots reach the crossing after ~70 ticks
    events = sc.spawn_events(ticks)
    dt_ns = int(round(float(sc.cfg.delta_t) * 1e9))
    for w in (g, o):
        w.set_environment_colliders(cols)
        w.set_tracking_buffers(capacity=16, sample_ns=100_000_000)
        if not obstacle_factors:
            w.change_factor_enabled(2, 0)
    totals = (0, 0)
    for tick in range(ticks):
        for _, k in [e for e in events if e[0] == tick]:
            sw = sc.spawn(k, rng)
            for w in (g, o):
                sw.add_to(w, set_sdf=False)
        if g.num_robots == 0:
            continue
        for w in (g, o):
            w.step()
        tg, to = g.update_environment_collisions(), o.update_environment_collisions()
        assert tg == to, f"{name} tick {tick}"
        totals = tg
        for w in (g, o):
            w.track(dt_ns, (tick + 1) * float(sc.cfg.delta_t))
    assert np.array_equal(g.read_environment_collisions(), o.read_environment_collisions())
    for r, (a, b) in enumerate(zip(g.read_tracks(), o.read_tracks())):
        for x, y in zip(a, b):
            assert x.shape == y.shape and np.array_equal(x, y), (name, r)
    assert g.num_robots >= 12 and len(g.read_tracks()[0][0]) >= 2
    if not obstacle_factors:
        assert totals[0] > 0, "the lanes of the junction pass the bollard"
