"""whole scenarios stepped on the device with the mission logic there too —
`reached_waypoint` (robot.rs:2080-2176) advances each robot's waypoint index from the estimated position
of the configured variable, the horizon prior then heads for the next waypoint and a robot that has passed
its last waypoint stops moving its horizon (FinishedPath).  Compared tick by tick with the oracle."""
import numpy as np
import pytest

from magics_b200 import World, scenarios
from magics_b200.sharded import LocalShards
from oracle.oracle import OracleWorld
from tests.test_gpu_parity import check

pytestmark = pytest.mark.gpu

CURRENT, HORIZON, VARIABLE = 0, 1, 2
RADIUS, METER = 0, 1


@pytest.mark.parametrize("make", [lambda cfg: World(cfg), lambda cfg: LocalShards(cfg, 3)], ids=["single", "ws3"])
def test_circle_runs_to_completion_on_device(make):
    sw = scenarios.circle(8, circle_radius=8.0)
    g, o = make(sw.cfg), OracleWorld(sw.cfg)
    sw.add_to(g)
    sw.add_to(o)
    task, fin = (HORIZON, 0, RADIUS, 0.0), (CURRENT, 0, RADIUS, 0.0)
    reached_total = np.zeros(sw.n, int)
    for tick in range(90):
        rg, ro = g.reached_waypoint(task, fin), o.reached_waypoint(task, fin)
        assert np.array_equal(rg, ro), f"tick {tick}: different robots reached a waypoint"
        reached_total += rg
        g.step()
        o.step()
        if tick % 15 == 0:
            check(g, o, f"mission tick {tick}")
    check(g, o, "mission end")
    assert np.array_equal(g.read_waypoint_index(), o.read_waypoint_index())
    assert (reached_total == 1).all() and (g.read_waypoint_index() == 2).all(), "every robot reaches its goal once"


def test_junction_three_waypoints_variable_and_meter_criteria():
    sw = scenarios.junction_twoway(per_lane=1)
    g, o = World(sw.cfg), OracleWorld(sw.cfg)
    sw.add_to(g)
    sw.add_to(o)
    task, fin = (VARIABLE, 4, METER, 6.0), (VARIABLE, 99, METER, 3.0)  # index 99 -> last variable
    seen = set()
    for tick in range(120):
        rg, ro = g.reached_waypoint(task, fin), o.reached_waypoint(task, fin)
        assert np.array_equal(rg, ro), f"tick {tick}"
        g.step()
        o.step()
        seen.update(g.read_waypoint_index().tolist())
        if tick % 20 == 0:
            check(g, o, f"junction mission tick {tick}")
    check(g, o, "junction mission end")
    assert np.array_equal(g.read_waypoint_index(), o.read_waypoint_index())
    assert {1, 2} <= seen, seen  # robots passed the junction-centre waypoint and headed for the exit


@pytest.mark.parametrize("make", [lambda cfg: World(cfg), lambda cfg: LocalShards(cfg, 3)], ids=["single", "ws3"])
@pytest.mark.parametrize("interrobot", [0, 1])
def test_robot_robot_collision_monitor(make, interrobot):
    """update_robot_robot_collisions (planner/collisions.rs:72-143).  With InterRobot
    factors disabled the circle swarm drives straight through its centre and collides; with them enabled
    it (mostly) does not.  Hit counts, currently-colliding pairs and per-robot counts against the oracle's
    all-pairs version, every tick."""
    sw = scenarios.circle(8, circle_radius=8.0, robot_radius=1.0)
    sw.cfg.enable_interrobot = interrobot
    g, o = make(sw.cfg), OracleWorld(sw.cfg)
    sw.add_to(g)
    sw.add_to(o)
    seen_now = 0
    for tick in range(60):
        g.step()
        o.step()
        cg, co = g.update_robot_collisions(), o.update_robot_collisions()
        assert cg == co, f"tick {tick}: (collisions, colliding now) {cg} vs {co}"
        seen_now = max(seen_now, cg[1])
        if tick % 10 == 9:
            assert np.array_equal(g.read_robot_collisions(), o.read_robot_collisions()), f"tick {tick}"
    check(g, o, "collision run")
    total, _ = g.update_robot_collisions()
    if not interrobot:
        assert total >= 4 and seen_now >= 2, (total, seen_now)
        assert g.read_robot_collisions().sum() == 2 * total


@pytest.mark.parametrize("make", [lambda cfg: World(cfg), lambda cfg: LocalShards(cfg, 3)], ids=["single", "ws3"])
def test_despawn_on_final_waypoint_and_mid_run_removal(make):
    """RobotDespawned (robot.rs:2171-2172, despawn_entity_after): a robot that completes its mission leaves the
    simulation; the next topology pass deletes every InterRobot factor to or from it (robot.rs:1386-1439, the
    failed `query.get_mut` branch) and nobody iterates or sees it again.  Two robots are also removed mid-run, while
    their InterRobot factors are active."""
    sw = scenarios.circle(10, circle_radius=9.0)
    g, o = make(sw.cfg), OracleWorld(sw.cfg)
    sw.add_to(g)
    sw.add_to(o)
    task, fin = (HORIZON, 0, RADIUS, 0.0), (CURRENT, 0, RADIUS, 0.0)
    gone = np.zeros(sw.n, bool)
    for tick in range(80):
        rg, ro = g.reached_waypoint(task, fin), o.reached_waypoint(task, fin)
        assert np.array_equal(rg, ro), f"tick {tick}: different robots reached a waypoint"
        done = (g.read_waypoint_index() >= 2) & ~gone  # mission.is_completed()
        if tick == 6:
            done[[2, 7]] = True
        if done.any():
            ids = np.flatnonzero(done).astype(np.int32)
            g.remove_robots(ids)
            o.remove_robots(ids)
            gone |= done
        g.step()
        o.step()
        if tick % 8 == 0 or done.any():
            check(g, o, f"despawn tick {tick}")
            off = g.read_connections()[0]
            assert (np.diff(off)[gone] == 0).all(), "a despawned robot keeps no connection"
            assert not np.isin(g.read_connections()[1], np.flatnonzero(gone)).any(), "nobody is connected to one"
    check(g, o, "despawn end")
    assert np.array_equal(g.read_removed().astype(bool), gone)
    assert gone.sum() >= 6, gone  # the two removed by hand and the robots that reached their goal
