"""The host side of the global-planner hand-off (planner/robot.rs:655-776): what `update_robot_mission` computes from an
arrived RRT* path before it calls `set_tracking_path` / `reset_variables` / `reset_tracking_factors` —
`magics_b200.mission.path_arrival` and `apply_global_paths`, checked against hand-evaluated f32 values and by running
the three calls on the CPU oracle (the device side of the same calls: tests/test_gpu_handoff.py)."""
import numpy as np
import pytest

from magics_b200 import scenarios
from magics_b200.mission import MissionClock, apply_global_paths, path_arrival
from oracle.oracle import OracleWorld


def test_waypoints_tracking_path_and_reset_means_as_written():
    f = np.float32
    wps, track, means = path_arrival([(0, 0), (10, 0), (10, 5)], target_speed=5.0, planning_horizon=5.0, num_variables=12)
    # velocity part = speed * normalize(from - to): it points BACK along the path (robot.rs:658), zero on the last point
    assert wps.tolist() == [[0, 0, -5, 0], [10, 0, 0, -5], [10, 5, 0, 0]] and wps.dtype == np.float32
    assert track.tolist() == [[0, 0], [10, 0], [10, 5]]
    # start / next as 4-vectors: dir = (10, 0, 5, -5), |dir| = sqrt(150); s = min(25, 0.9 |dir|)
    d = np.array([10, 0, 5, -5], f)
    length = np.sqrt(f(150.0))
    s = f(length * f(0.9))
    nxt = f(0) + s * (d / length)
    assert means.shape == (12, 4) and means.dtype == np.float64
    for i in range(12):
        r = f(i) / f(12)
        assert means[i, 0] == float(f(0) + (nxt[0] - f(0)) * r) and means[i, 1] == 0.0
        assert means[i, 2] == float(f(5) * (d / length)[0]) and means[i, 3] == float(f(5) * (d / length)[1])
    # a long first segment: the horizon (speed x planning horizon) caps the spread instead of 0.9 |dir|
    _, _, far = path_arrival([(0, 0), (100, 0)], 2.0, 5.0, 10)
    assert far[-1, 0] == pytest.approx(0.9 * 10 * 100 / np.sqrt(100 ** 2 + 2 ** 2), rel=1e-6)  # next' = 10 along dir; i / V = 0.9
    with pytest.raises(ValueError):
        path_arrival([(1, 1)], 1.0, 1.0, 4)


def test_hand_off_on_the_oracle_resets_and_the_robot_follows_the_new_path():
    sw = scenarios.junction_twoway(per_lane=1)
    o = OracleWorld(sw.cfg)
    sw.add_to(o)
    clock = MissionClock()
    clock.spawn([sw.wp_xy[sw.wp_offsets[r]:sw.wp_offsets[r + 1]] for r in range(sw.n)], started_at=0.0)
    for _ in range(3):
        o.step()
    robots = [0, 3]
    pos = o.read_positions()
    paths = [np.array([pos[r], pos[r] + np.array([8.0, 3.0], np.float32), pos[r] + np.array([16.0, -2.0], np.float32)],
                      np.float32) for r in robots]
    apply_global_paths(o, robots, paths, sw.cfg.target_speed, 5.0, clock)
    b = o.read_beliefs()
    for r, p in zip(robots, paths):
        _, track, means = path_arrival(p, sw.cfg.target_speed, 5.0, sw.cfg.num_variables)
        assert np.array_equal(b["mean"][r], means)  # reset_variables replaced every mean (factorgraph.rs:1541-1564)
        assert clock.missions[r].route.target_index == 1 and len(clock.missions[r].route.waypoints) == 3
    assert (o.read_waypoint_index()[robots] == 1).all()
    before = [np.linalg.norm(o.read_positions()[r] - paths[k][1]) for k, r in enumerate(robots)]
    for _ in range(25):
        o.step()
    after = [np.linalg.norm(o.read_positions()[r] - paths[k][1]) for k, r in enumerate(robots)]
    assert np.isfinite(o.read_beliefs()["mean"]).all()
    assert all(a < b0 - 2.0 for a, b0 in zip(after, before)), (before, after)  # heading for the path's second point
