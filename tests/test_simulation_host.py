"""`magics_b200.simulation.Simulation` — the headless runner of a reference scenario directory — driven without a GPU:
the world it steps is the CPU oracle behind the method surface of `magics_b200.World`.  Checks the order of the systems
around the iteration (spawner clock, mission clocks, despawn, antenna draws, collision monitors + entry lists, trackers)
and that what comes out is the reference's ExportData, readable by `magics_b200.metrics`."""
import json
import os

import numpy as np
import pytest

from magics_b200 import metrics
from magics_b200.scenarios import ReferenceScenario
from magics_b200.simulation import Simulation
from oracle import oracle
from oracle.oracle import OracleWorld


class OracleAsWorld(OracleWorld):
    """The oracle with the two members of `magics_b200.World` it lacks (`cfg`, `export_totals`)."""

    def __init__(self, cfg, env):
        super().__init__(cfg, threads=4)
        self.cfg = cfg
        self.set_sdf(oracle.env_to_sdf_image(env))
        self._has_colliders = False

    def set_environment_colliders(self, colliders):
        self._has_colliders = True
        return super().set_environment_colliders(colliders)

    def export_totals(self):
        n = self.num_robots
        if n:
            self.update_robot_collisions()  # refreshes the per-robot counts the wrapper caches (state unchanged)
        return {"collisions_robots": self.read_robot_collisions() if n else np.zeros(0, np.uint32),
                "next_waypoint": self.read_waypoint_index(), "removed": self.read_removed().astype(bool),
                "collisions_environment": self.read_environment_collisions() if self._has_colliders and n else None,
                "tracks": self.read_tracks() if n else None, "messages": None}


def _sim(name, **kw):
    sc = ReferenceScenario(name)  # tests/golden/scenarios.json: the reference's own files, extracted
    return sc, Simulation(sc, OracleAsWorld(sc.cfg, sc.env), np.random.default_rng(4), **kw)


def test_structured_junction_twoway_runs_spawns_despawns_and_exports():
    sc, sim = _sim("Structured Junction Twoway")
    assert sim.dt_ns == 100_000_000 and len(sim.colliders) == 4
    sim.run(ticks=1)
    assert sim.world.num_robots == 4  # the four lanes with delay 0; the others follow after 2 s and 4 s
    sim.run(ticks=230)  # every lane repeats after 6 s; the first robots cross the junction and leave after ~16 s
    n = sim.world.num_robots
    assert n == 48 and len(sim.clock.missions) == n and sim.radii.shape == (n,)
    assert sim.gone.sum() >= 6 and np.array_equal(sim.gone, sim.world.read_removed().astype(bool))
    assert all(m.completed for m, g in zip(sim.clock.missions, sim.gone) if g)
    d = json.loads(json.dumps(sim.export()))
    assert d["scenario"] == "Structured Junction Twoway" and d["makespan"] == pytest.approx(23.1)
    assert d["delta_t"] == pytest.approx(0.1) and set(d["collisions"]) == {"robots", "environment"}
    assert len(d["robots"]) == n and len(d["obstacles"]) == 4 and d["goal_areas"] == {}
    first = d["robots"]["0"]
    assert first["mission"]["started_at"] == 0.0 and first["mission"]["finished_at"] > 5.0
    assert d["robots"]["12"]["mission"]["started_at"] == pytest.approx(6.0)  # the second wave
    assert 100 <= len(first["positions"]) <= 231  # 100 ms tracker: one sample per tick while it moves
    ev = metrics.evaluate(d, projection="segments")
    done = [rid for rid, g in enumerate(sim.gone) if g]
    assert all(40.0 < ev["robots"][str(r)]["distance_travelled"] < 140.0 for r in done)
    assert ev["collision_entries"] == {"interrobot": len(sim.log.robot_entries),
                                       "environment": len(sim.log.environment_entries)}


def _run_reference_script(script, *args):
    import subprocess
    import sys

    runner = (
        "import sys, runpy\n"
        "from unittest import mock\n"
        "for name in ('matplotlib', 'matplotlib.pyplot', 'toolz', 'toolz.curried', 'seaborn', 'result'):\n"
        "    sys.modules[name] = mock.MagicMock(name=name)\n"  # plotting only; not in this image
        "script = sys.argv[1]; sys.argv = sys.argv[1:]\n"
        "sys.path.insert(0, '/root/reference/scripts')\n"
        "runpy.run_path(script, run_name='__main__')\n")
    return subprocess.run([sys.executable, "-c", runner, f"/root/reference/scripts/{script}", *args],
                          capture_output=True, text=True, timeout=120, env={**os.environ, "COLUMNS": "200"})


@pytest.mark.skipif(not os.path.isdir("/root/reference/scripts"), reason="reference tree only in the build container")
def test_reference_scripts_read_what_the_runner_exports(tmp_path):
    """The INTEGRATION.md recipe end to end (minus the GPU): run a scenario, export, feed the reference's own scripts."""
    sc, sim = _sim("Structured Junction Twoway")
    sim.run(ticks=60)
    path = tmp_path / "export_structured junction twoway_0.json"
    path.write_text(json.dumps(sim.export()))
    # (perpendicular-path-deviation.py divides by x2 - x1 of a route segment, :41, and trips over its own NaN on the
    # exactly vertical lanes of this formation; tests/test_metrics_host.py runs it on routes it can handle)
    for script, args in (("ldj.py", ("-i", str(path))), ("distance-travelled.py", (str(path),))):
        p = _run_reference_script(script, *args)
        assert p.returncode == 0, (script, p.stderr[-1500:])
        assert len(p.stdout) > 100, script


def test_run_until_every_mission_is_complete():
    sc, sim = _sim("Circle Experiment", environment_collisions=False)
    sc.formations[0].robots = 4  # the experiment sweeps 5 ... 50 robots; four keep the CPU oracle quick (V = 21, 50 / 10 iterations)
    steps = sim.run(max_ticks=400)
    assert sim.world.num_robots == 4 and all(m.completed for m in sim.clock.missions)
    assert 30 < steps < 400, steps  # 100 m at 15 m / s and 10 Hz: about 70 ticks
    d = sim.export()
    assert d["makespan"] == pytest.approx(steps * 0.1) and len(d["collisions"]["robots"]) == 0
    finished = [r["mission"]["finished_at"] for r in d["robots"].values()]
    assert max(finished) <= d["makespan"] + 1e-9 and min(finished) > 3.0


def test_antenna_draws_reach_the_world_when_the_scenario_has_a_failure_rate():
    sc, sim = _sim("Circle Experiment", environment_collisions=False)
    sc.formations[0].robots = 6
    sc.failure_rate = 0.5
    calls = []
    real = sim.world.set_comms
    sim.world.set_comms = lambda antenna_active=None, idle=None: (calls.append(np.array(antenna_active)), real(antenna_active, idle))[1]
    first = sim.scenario.spawn_events(200)[0][0]  # the formation's delay
    sim.run(ticks=first + 6)
    assert len(calls) == 6 and all(c.shape == (6,) for c in calls) and 0 < np.concatenate(calls).mean() < 1


@pytest.mark.skipif(not os.path.isdir("/root/reference/config/scenarios"), reason="reference tree only in the build container")
def test_command_line_runner_refuses_to_run_without_a_gpu(tmp_path):
    """`python -m magics_b200.simulation <dir>` parses the scenario directory, then needs the engine: no CPU fallback."""
    import subprocess
    import sys

    from tests.conftest import _cuda_devices

    if _cuda_devices() > 0:
        pytest.skip("a GPU is present")
    p = subprocess.run([sys.executable, "-m", "magics_b200.simulation", "/root/reference/config/scenarios/Junction Twoway",
                        "--ticks", "3", "--export", str(tmp_path / "o.json")], capture_output=True, text=True, timeout=120,
                       cwd=os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    assert p.returncode != 0 and "CUDA" in p.stderr and not (tmp_path / "o.json").exists()


@pytest.mark.skipif(not os.path.isdir("/root/reference/config/scenarios"), reason="reference tree only in the build container")
def test_rrt_star_formations_wait_for_the_planner_follow_its_path_and_hand_over_route_by_route():
    """`Solo GP` (planning-strategy: rrt-star, tracking factors on) with a second waypoint added, so that the mission has
    three taskpoints = two routes.  The planner is a stub returning a dog-leg (the RRT* search is outside this repo);
    what is checked is progress_missions' state machine around it (robot.rs:562-812) and the hand-off calls."""
    import copy

    sc = ReferenceScenario.from_directory("/root/reference/config/scenarios/Solo GP")
    f = sc.formations[0]
    assert f.planning_strategy == "rrt-star" and sc.cfg.enable_tracking == 1
    f.waypoints.append(copy.deepcopy(f.waypoints[0]))  # taskpoint 2: back where taskpoint 1's shape puts it, shifted below
    asked = []

    def planner(start, end, colliders, rng):
        asked.append((tuple(start), tuple(end), len(colliders)))
        if len(asked) == 2:
            return None  # a failed search: the mission goes back to Idle and asks again
        mid = (0.5 * (start[0] + end[0]) + 6.0, 0.5 * (start[1] + end[1]) - 6.0)
        return [start, mid, end]

    world = OracleAsWorld(sc.cfg, sc.env)
    world.change_factor_enabled(2, 0)  # the stub's dog-leg cuts through walls: Obstacle factors off, or the robot stops at one
    idle_seen = []
    real = world.set_comms
    world.set_comms = lambda antenna_active=None, idle=None: (idle_seen.append(None if idle is None else int(idle[0])),
                                                              real(antenna_active, idle))[1]
    sim = Simulation(sc, world, np.random.default_rng(2), global_planner=planner)
    first = sc.spawn_events(100)[0][0]
    sim.run(ticks=first + 1)
    m = sim.clock.missions[0]
    # identical second waypoint -> the route from taskpoint 1 to 2 is degenerate; move taskpoint 2 so the test has a second leg
    m.taskpoints[2] = (m.taskpoints[1][0] - 40.0, m.taskpoints[1][1])
    assert len(m.taskpoints) == 3 and m.state == "waiting" and idle_seen[-1] == 1 and len(asked) == 1
    start0 = sim.world.read_positions()[0].copy()
    assert np.array_equal(start0, np.asarray(m.taskpoints[0], np.float32)), "an idle robot does not move (robot.rs:2303)"
    sim.run(ticks=1)
    assert m.state == "active" and len(m.route.waypoints) == 3 and m.route.target_index == 1  # the dog-leg replaced the 2 taskpoints
    assert idle_seen[-1] == 0 and not np.array_equal(sim.world.read_positions()[0], start0)  # Active: it iterates and moves
    assert int(sim.world.read_waypoint_index()[0]) == 1
    states = []
    for _ in range(1500):
        sim.tick()
        states.append(m.state)
        if m.completed:
            break
    assert m.completed and m.finished_at is not None and len(m.routes) == 2
    # seen once per tick: route 1 active -> (its last waypoint is reached: next_route, Idle, the planner is asked in the same
    # pass) waiting -> (failed search) idle -> waiting -> active -> completed
    order = [s for k, s in enumerate(states) if k == 0 or s != states[k - 1]]
    assert order == ["active", "waiting", "idle", "waiting", "active", "completed"], order
    assert len(asked) == 3 and asked[1][0] == pytest.approx(m.taskpoints[1]) and asked[1][2] == len(sim.colliders) > 0
    assert m.routes[0].finished_at is not None and m.routes[1].started_at >= m.routes[0].started_at
    d = sim.export()
    md = d["robots"]["0"]["mission"]
    assert len(md["routes"]) == 2 and len(md["waypoints"]) == 3 and all(len(r["waypoints"]) == 3 for r in md["routes"])
    assert sim.gone[0] == sc.despawn


def test_formations_with_different_waypoint_criteria_each_get_their_own():
    """Two formations of one scenario with different `finished-when-intersects` (as in `Collaborative GP`): the engine's
    reached_waypoint takes one criterion pair for all robots, the runner calls it per pair and keeps, per robot, the
    outcome of the robot's own formation — equal to running each formation alone."""
    import copy

    def scenario(which):
        sc = ReferenceScenario("Structured Junction Twoway")
        a, b = copy.deepcopy(sc.formations[0]), copy.deepcopy(sc.formations[3])  # two lanes that spawn at t = 0
        b.finished_when = (0, 0, 1, 25.0)  # `current` within 25 m: finishes far earlier than lane a's criterion
        assert a.finished_when != b.finished_when
        sc.formations = [f for f, keep in ((a, "a" in which), (b, "b" in which)) if keep]
        for f in sc.formations:
            f.repeat_every_s, f.repeat_times = None, None  # one robot each
            f.placement = "equal"  # no random draw: the robot of a lane starts at the same point in every run
        return sc

    def run(which, ticks=260):
        sc = scenario(which)
        world = OracleAsWorld(sc.cfg, sc.env)
        world.change_factor_enabled(1, 0)  # no InterRobot factors: the two robots do not influence each other
        sim = Simulation(sc, world, np.random.default_rng(3), environment_collisions=False)
        sim.run(ticks=ticks)
        return [(m.route.target_index, m.finished_at) for m in sim.clock.missions]

    both, only_a, only_b = run("ab"), run("a"), run("b")
    assert both == only_a + only_b
    assert both[0][1] is not None and both[1][1] is not None and both[1][1] < both[0][1] - 1.0


_REF = "/root/reference/config/scenarios"


@pytest.mark.skipif(not os.path.isdir(_REF), reason="reference tree only in the build container")
@pytest.mark.parametrize("name", sorted(os.listdir(_REF)) if os.path.isdir(_REF) else [])
def test_every_shipped_scenario_runs_through_the_runner_and_exports(name):
    """All 18 scenario directories of the reference, 25 fixed steps past their first spawn on the oracle (swarm sizes cut
    where the oracle would be slow: V up to 35, 50 internal iterations), colliders from the scenario's own environment,
    export serialisable."""
    sc = ReferenceScenario.from_directory(os.path.join(_REF, name))
    for f in sc.formations:
        f.robots = min(f.robots, 4 if (sc.cfg.num_variables > 21 or sc.cfg.iterations_internal > 20) else 8)
    sim = Simulation(sc, OracleAsWorld(sc.cfg, sc.env), np.random.default_rng(0))
    events = sc.spawn_events(600)
    sim.run(ticks=(events[0][0] if events else 0) + 25)
    d = json.loads(json.dumps(sim.export()))
    assert d["scenario"] == name and len(d["robots"]) == sim.world.num_robots == len(sim.clock.missions)
    assert len(d["obstacles"]) == len(sim.colliders) and set(d["collisions"]) == {"robots", "environment"}
    if sim.world.num_robots:
        assert all(len(r["positions"]) >= 1 or r["mission"]["routes"] for r in d["robots"].values())
        assert np.isfinite(sim.world.read_beliefs()["mean"]).all()
    else:
        assert name == "Obstacle Shapes Showcase"
