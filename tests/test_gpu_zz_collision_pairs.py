"""The export's collision entry lists with the ENGINE as the world: `CollisionLog` over `magics_b200.World` must produce
the entries, box for box, that the oracle records itself (planner/collisions.rs:117-138, :417-426, :700-716;
export.rs:171-214, :552-555).  The log's logic is checked without a GPU in tests/test_collision_log_host.py (despawn and
mid-run spawn under the robot-robot monitor included); the two runs here are the scenarios of the monitor tests in
tests/test_gpu_mission.py / tests/test_gpu_evaluation.py.  Verified on a B200: profiles/r02w."""
import json

import numpy as np
import pytest

from magics_b200 import World, scenarios
from magics_b200.collisions import CollisionLog
from magics_b200.environment import Collider, Environment, tile_colliders
from magics_b200.export import export_data
from oracle.oracle import OracleWorld
from tests.test_collision_log_host import SHAPES, assert_same_entries

pytestmark = pytest.mark.gpu


def test_robot_robot_entries_of_the_engine_equal_the_oracles():
    sw = scenarios.circle(8, circle_radius=8.0, robot_radius=1.0)
    sw.cfg.enable_interrobot = 0
    g, o = World(sw.cfg), OracleWorld(sw.cfg)
    sw.add_to(g)
    sw.add_to(o)
    log = CollisionLog(g, sw.radii)
    for tick in range(70):
        g.step()
        o.step()
        assert log.update_robot_collisions() == o.update_robot_collisions(), f"tick {tick}"
    assert_same_entries(log.robot_entries, o.read_collision_events(0), "robot-robot")
    assert len(log.robot_entries) >= 4


def test_robot_environment_entries_of_the_engine_equal_the_oracles_and_reach_the_export():
    env = Environment(grid=["┼"], tile_size=100.0, path_width=0.1325)
    cols = tile_colliders(env) + [Collider("ball", (0.0, 0.0), 0.0, radius=2.0)] + SHAPES[3:]
    sw = scenarios.circle(12, 45.0, robot_radius=1.0)
    g, o = World(sw.cfg), OracleWorld(sw.cfg)
    sw.add_to(g)
    sw.add_to(o)
    for w in (g, o):
        w.change_factor_enabled(2, 0)
        w.set_environment_colliders(cols)
    log = CollisionLog(g, sw.radii, cols)
    for tick in range(140):
        g.step()
        o.step()
        assert log.update_environment_collisions() == o.update_environment_collisions(), f"tick {tick}"
        if tick == 30:
            for w in (g, o):
                sw.add_to(w, set_sdf=False)
            log.add_robots(sw.radii)
        if tick == 60:
            for w in (g, o):
                w.remove_robots([1, 13])
    assert_same_entries(log.environment_entries, o.read_collision_events(1), "robot-environment")
    d = json.loads(json.dumps(export_data(g, scenario="junction", radii=log.radii, colliders=cols, collision_log=log)))
    assert len(d["collisions"]["environment"]) == len(log.environment_entries) >= 12
    for r in range(g.num_robots):
        mine = sum(len(e["aabbs"]) for e in d["collisions"]["environment"] if e["robot"] == r)
        assert mine == d["robots"][str(r)]["collisions"]["environment"]
