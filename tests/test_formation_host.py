"""Scenario front-end on the CPU: `Formation::as_positions` and the FormationSpawner clock
restated in magics_b200/formation.py, run on the reference's own scenario inputs (tests/golden/scenarios.json, extracted
by tests/golden/make_golden.py from config/scenarios/*/).  The reference ships no test for these functions, so they are
checked against independent float64 geometry and against the invariants the code promises (no overlap, order)."""
import json
import os

import numpy as np
import pytest

from magics_b200 import formation as F
from magics_b200 import scenarios
from oracle import oracle


@pytest.fixture(autouse=True)
def _timesteps_from_the_oracle(monkeypatch):
    # host-only test: no product library needed for the variable timesteps
    monkeypatch.setattr(scenarios, "get_variable_timesteps", oracle.variable_timesteps)


def _golden(golden_dir):
    return json.load(open(os.path.join(golden_dir, "scenarios.json")))


def test_golden_scenarios_parse(golden_dir):
    d = _golden(golden_dir)
    assert set(d) == {"Circle Experiment", "Structured Junction Twoway", "Collaborative Complex"}
    counts = {k: len(v["formations"]) for k, v in d.items()}
    assert counts == {"Circle Experiment": 1, "Structured Junction Twoway": 12, "Collaborative Complex": 19}
    for name in d:
        sc = scenarios.ReferenceScenario(name)
        assert sc.reached_when is not None, "one criterion pair per formation group"
        for f in sc.formations:
            assert f.robots >= 1 and len(f.waypoints) >= 1
    sc = scenarios.ReferenceScenario("Structured Junction Twoway")
    assert sc.cfg.num_variables == 12 and sc.cfg.enable_tracking == 1  # BASELINE config 2: V = 12, tracking on
    assert sc.finished_when == (F.VARIABLE, 5, F.METER, 10.0) and sc.reached_when == (F.HORIZON, 0, F.ROBOT_RADIUS, 0.0)
    sc = scenarios.ReferenceScenario("Collaborative Complex")
    assert (sc.world_w, sc.world_h) == (250.0, 175.0) and sc.env.image_shape == (1400, 2000)  # config 3: SDF 2000x1400
    sc = scenarios.ReferenceScenario("Circle Experiment")
    assert sc.cfg.num_variables == 21 and sc.cfg.iterations_internal == 50  # target-speed 15: V = 21


def test_circle_formation_equal_placement_and_cross_projection():
    sc = scenarios.ReferenceScenario("Circle Experiment")
    f = sc.formations[0]
    radii = np.full(f.robots, 2.5, np.float32)
    init, wps = F.as_positions(f, sc.world_w, sc.world_h, radii, np.random.default_rng(0))
    assert init.dtype == np.float32 and init.shape == (30, 2) and len(wps) == 1
    ang = np.arange(30) * (2 * np.pi / 30)
    ref = 50.0 * np.stack([np.cos(ang), np.sin(ang)], axis=1)  # centre (0.5, 0.5) -> world origin
    assert np.abs(init - ref).max() < 2e-5
    assert np.abs(wps[0] + ref).max() < 1e-4  # Cross: the antipode (f32 angle + pi, f32 cos / sin at radius 50)
    # the first robot sits exactly on the x axis; its antipode carries f32 sin(pi) != 0, as in the reference
    assert init[0].tolist() == [50.0, 0.0] and wps[0][0, 1] != 0.0
    r = F.routes(init, wps)
    assert len(r) == 30 and r[3].shape == (2, 2) and np.array_equal(r[3][0], init[3]) and np.array_equal(r[3][1], wps[0][3])


def test_line_segment_random_placement_keeps_robots_apart_and_projects_in_order():
    f = F.Formation(robots=5, delay_s=0.0, repeat_every_s=None, repeat_times=None,
                    initial_shape=F.ShapeSpec("line-segment", points=((0.1, 0.2), (0.1, 0.8))), placement="random",
                    attempts=500, waypoints=[(F.ShapeSpec("line-segment", points=((0.9, 0.2), (0.9, 0.8))), "identity"),
                                             (F.ShapeSpec("line-segment", points=((0.5, 0.0), (0.5, 1.0))), "cross")])
    radii = np.array([1.0, 2.0, 1.5, 1.0, 3.0], np.float32)
    init, wps = F.as_positions(f, 100.0, 100.0, radii, np.random.default_rng(7))
    assert np.all(init[:, 0] == np.float32(-40.0))  # ((0.1 - 0.5) * 100) as f32
    d = np.abs(init[:, None, 1] - init[None, :, 1])
    need = radii[:, None] + radii[None, :]
    iu = np.triu_indices(5, 1)
    assert np.all(d[iu] >= need[iu] - 1e-5)
    # identity keeps every robot's lerp amount, cross hands them out in reverse robot order
    lerp = (init[:, 1] + 30.0) / 60.0
    assert np.abs(wps[0][:, 1] - (-30.0 + 60.0 * lerp)).max() < 1e-4
    assert np.abs(wps[1][:, 1] - (-50.0 + 100.0 * lerp[::-1])).max() < 1e-4
    # a segment too short for the robots: None after `attempts` tries (the reference logs and skips the spawn)
    f.initial_shape = F.ShapeSpec("line-segment", points=((0.1, 0.5), (0.1, 0.51)))
    assert F.as_positions(f, 100.0, 100.0, radii, np.random.default_rng(7)) is None


def test_line_segment_equal_placement_follows_the_code_as_written():
    f = F.Formation(robots=3, delay_s=0.0, repeat_every_s=None, repeat_times=None,
                    initial_shape=F.ShapeSpec("line-segment", points=((0.2, 0.5), (0.8, 0.5))), placement="equal",
                    attempts=0, waypoints=[(F.ShapeSpec("line-segment", points=((0.2, 0.9), (0.8, 0.9))), "identity")])
    radii = np.array([1.0, 1.0, 1.0], np.float32)
    init, wps = F.as_positions(f, 100.0, 100.0, radii, None)
    # first centre one radius in; lerp amounts strictly increasing; same amounts on the waypoint segment
    assert init[0].tolist() == [-29.0, 0.0]
    lerp = (init[:, 0] + 30.0) / 60.0
    assert np.all(np.diff(lerp) > 0)
    assert np.abs(wps[0][:, 0] - (-30.0 + 60.0 * lerp)).max() < 1e-4
    # evenly_place...: None when length / max_radius < min_radius
    f.initial_shape = F.ShapeSpec("line-segment", points=((0.5, 0.5), (0.505, 0.5)))
    assert F.as_positions(f, 100.0, 100.0, np.array([2.0, 2.0, 2.0], np.float32), None) is None


def test_spawner_clock():
    sc = scenarios.ReferenceScenario("Structured Junction Twoway")
    ev = sc.spawn_events(130)
    # delays 0 / 2 / 4 s, every 6 s at 10 Hz: four formations per wave
    by_tick = {}
    for t, k in ev:
        by_tick.setdefault(t, []).append(k)
    assert sorted(by_tick) == [0, 20, 40, 60, 80, 100, 120]
    assert all(len(v) == 4 for v in by_tick.values())
    assert by_tick[0] == by_tick[60] == by_tick[120]
    # `times: !finite n` = n spawns in all (first + n - 1 repeats); no `repeat`: exactly one
    f = sc.formations[0]
    assert len(f.spawn_times(1e9)) == 100
    c = scenarios.ReferenceScenario("Circle Experiment").formations[0]
    assert c.spawn_times(1e9) == [1.0]


def test_spawn_builds_routes_and_initial_means():
    sc = scenarios.ReferenceScenario("Collaborative Complex")
    sw = sc.spawn(1, np.random.default_rng(3))
    assert sw.n == 1 and sw.init_means.shape == (1, sc.cfg.num_variables, 4)
    wp = sw.wp_xy.reshape(-1, 2)
    assert wp.shape[0] == len(sc.formations[1].waypoints) + 1 and np.array_equal(wp[0], sw.positions[0])
    # initial velocity: target speed toward the first waypoint (spawner.rs:470-482)
    d = wp[1] - wp[0]
    v = sw.init_means[0, 0, 2:]
    assert np.allclose(v, sc.cfg.target_speed * d / np.linalg.norm(d), atol=1e-5)
    # the horizon variable sits planning_horizon * target_speed ahead (robot.rs:1157-1192)
    ahead = np.linalg.norm(sw.init_means[0, -1, :2] - sw.init_means[0, 0, :2])
    assert abs(ahead - min(np.linalg.norm(d), sc.planning_horizon * sc.cfg.target_speed)) < 1e-3
