"""Global-planner hand-off on the device (planner/robot.rs:655-776): new tracking path,
FactorGraph::reset_variables, reset_tracking_factors — parity with the oracle's literal restatement,
beliefs within 1e-9 (measured 0.0), indexing exact."""
import numpy as np
import pytest

from magics_b200 import World, scenarios
from magics_b200.sharded import LocalShards
from oracle.oracle import OracleWorld
from tests.test_gpu_parity import check, make_pair

pytestmark = pytest.mark.gpu


def _handoff_means(V, start, nxt, speed):
    """The means update_robot_mission resets to (robot.rs:739-758): positions lerped start -> next by i / n,
    velocity = speed * direction (f32 like the reference's Vec4 arithmetic, widened)."""
    start, nxt = np.asarray(start, np.float32), np.asarray(nxt, np.float32)
    d = nxt - start
    dn = d / np.linalg.norm(d)
    out = np.zeros((V, 4))
    for i in range(V):
        r = np.float32(i) / np.float32(V)
        out[i, :2] = start + (nxt - start) * r
        out[i, 2:] = np.float32(speed) * dn
    return out


def test_rrt_path_arrives_for_some_robots_of_the_junction():
    sw = scenarios.junction_twoway(per_lane=2)
    g, o = make_pair(sw)
    V = sw.cfg.num_variables
    for _ in range(3):  # tracking factors are live from the second tick on (20 factor iterations per tick)
        g.step()
        o.step()
    check(g, o, "before hand-off")
    robots = [0, 5, 6, 17]
    pos = o.read_positions()
    paths, means = [], []
    for k, r in enumerate(robots):
        p0 = pos[r]
        p1 = p0 + np.array([6.0 + k, -4.0 + 2 * k], np.float32)
        p2 = p1 + np.array([10.0, 3.0], np.float32)
        paths.append(np.array([p0, p1, p2], np.float32))
        means.append(_handoff_means(V, p0, p1, sw.cfg.target_speed))
    for w in (g, o):
        w.set_tracking_path(robots, paths)
        w.reset_variables(robots, np.array(means))
        w.reset_tracking_factors(robots)
    check(g, o, "right after hand-off")  # lam = diag(inf) on the interior variables, means replaced
    assert np.array_equal(g.read_waypoint_index(), o.read_waypoint_index())
    for tick in range(6):
        g.step()
        o.step()
        check(g, o, f"tick {tick} after hand-off")


@pytest.mark.parametrize("ws", [1, 3])
def test_reset_of_robots_with_active_interrobot_factors(ws):
    """Both ends of an edge reset, one end reset, and a reset while the neighbour's radio is off: the robot's own
    InterRobot factors lose the neighbour's message until the neighbour delivers again.  ws = 3: the swarm is split
    over three shards, so most reset robots have neighbours whose factors live on another shard."""
    sw = scenarios.circle(12, circle_radius=14.0)
    if ws == 1:
        g, o = make_pair(sw)
    else:
        g, o = LocalShards(sw.cfg, ws), OracleWorld(sw.cfg)
        sw.add_to(g)
        sw.add_to(o)
    V = sw.cfg.num_variables
    for _ in range(4):
        g.step()
        o.step()
    rng = np.random.default_rng(2)
    for rnd, robots in enumerate(([1, 2], [7], [0, 4, 5, 11])):
        b = o.read_beliefs()
        means = b["mean"].reshape(12, V, 4)[robots] + rng.normal(0, 0.3, (len(robots), V, 4))
        ant = np.ones(12, np.uint8)
        if rnd == 1:
            ant[[6, 8]] = 0  # the reset robot's neighbours cannot deliver for two ticks
        for w in (g, o):
            w.reset_variables(robots, means)
            w.set_comms(ant, None)
        check(g, o, f"round {rnd}: after reset")
        for tick in range(3):
            if tick == 2:
                for w in (g, o):
                    w.set_comms(np.ones(12, np.uint8), None)
            g.step()
            o.step()
            check(g, o, f"round {rnd} tick {tick}")


def test_half_steps_right_after_a_reset():
    sw = scenarios.circle(8, circle_radius=10.0)
    g, o = make_pair(sw)
    V = sw.cfg.num_variables
    for w in (g, o):
        w.step()
        w.step()
    means = o.read_beliefs()["mean"].reshape(8, V, 4)[[3]] * 0.9
    for w in (g, o):
        w.reset_variables([3], means, first_last_sigma=1e30, inbetween_sigma=float("inf"))
    for half in ("ext", "int", "ext", "int", "int", "ext"):
        for w in (g, o):
            if half == "ext":
                w.external_factor_iteration()
                w.external_variable_iteration()
            else:
                w.internal_factor_iteration()
                w.internal_variable_iteration()
        check(g, o, f"half {half} after reset", connectivity=False)


def test_tracking_path_only_and_errors():
    sw = scenarios.junction_twoway(per_lane=1)
    g, o = make_pair(sw)
    for w in (g, o):
        w.step()
        w.step()
    pos = o.read_positions()
    path = np.array([pos[2], pos[2] + [5, 5], pos[2] + [12, 4], pos[2] + [20, 9]], np.float32)
    for w in (g, o):
        w.set_tracking_path([2], [path])
        w.reset_tracking_factors([2, 4])
    for tick in range(12):  # past the 10-iteration timeout
        g.step()
        o.step()
    check(g, o, "new path + timeout")
    with pytest.raises(RuntimeError):
        g.set_tracking_path([1], [path[:1]])
    with pytest.raises(RuntimeError):
        g.reset_variables([99], np.zeros((1, sw.cfg.num_variables, 4)))
