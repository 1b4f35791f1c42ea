"""`magics_b200.planner` — RRT* over the reference's `CollisionProblem` (gbp_global_planner/src/lib.rs:135-187) for the
`planning-strategy: rrt-star` formations — on the reference's own `Solo GP` scenario, without a GPU: the path is
collision free for a Ball(collision-radius), and the headless runner takes the robot along it to its goal on the CPU
oracle with every factor kind of the scenario enabled (Dynamic, Obstacle on the generated SDF, Tracking on the planned
path).  The search is this repo's own (the `rrt` crate is third party and random): no parity claim, see the module."""
import os
import tomllib

import numpy as np
import pytest

from magics_b200.environment import Collider, colliders as environment_colliders
from magics_b200.planner import Feasibility, RRTStarPlanner, rrt_star, smooth_path
from magics_b200.scenarios import ReferenceScenario
from magics_b200.simulation import Simulation

SOLO = "/root/reference/config/scenarios/Solo GP"
needs_reference = pytest.mark.skipif(not os.path.isdir(SOLO), reason="reference tree only in the build container")


def test_feasibility_is_the_collision_monitors_predicate_and_rrt_star_goes_round_a_wall():
    wall = [Collider("cuboid", (0.0, 0.0), 0.0, half_extents=(1.0, 10.0))]
    f = Feasibility(wall, 1.0)
    assert f([(-5.0, 0.0), (0.0, 0.0), (1.9, 0.0), (2.1, 0.0), (0.0, 11.5)]).tolist() == [True, False, False, True, True]
    assert not f.segment((-5.0, 0.0), (5.0, 0.0), 0.5) and f.segment((-5.0, 12.0), (5.0, 12.0), 0.5)
    rng = np.random.default_rng(0)
    path = rrt_star((-8.0, 0.0), (8.0, 0.0), f, rng, step_size=2.0, neighbourhood_radius=4.0, max_iterations=20000,
                    bounds=((-15.0, -15.0), (15.0, 15.0)))
    assert path is not None and path[0] == (-8.0, 0.0) and path[-1] == (8.0, 0.0)
    assert all(f.segment(a, b, 0.25) for a, b in zip(path, path[1:]))
    assert max(abs(p[1]) for p in path) > 10.0  # round the end of the wall
    short = smooth_path(path, f, rng, step_size=0.25, max_iterations=300)
    length = lambda p: sum(np.hypot(b[0] - a[0], b[1] - a[1]) for a, b in zip(p, p[1:]))
    assert short[0] == path[0] and short[-1] == path[-1] and length(short) <= length(path) + 1e-9
    assert all(f.segment(a, b, 0.25) for a, b in zip(short, short[1:]))
    # a goal walled in on every side: the budget runs out -> None (PathfindingError::ReachedMaxIterations)
    box = [Collider("cuboid", (8.0, 0.0), 0.0, half_extents=(4.0, 0.5)), Collider("cuboid", (8.0, 6.0), 0.0, half_extents=(4.0, 0.5)),
           Collider("cuboid", (4.0, 3.0), 0.0, half_extents=(0.5, 3.5)), Collider("cuboid", (12.0, 3.0), 0.0, half_extents=(0.5, 3.5))]
    assert rrt_star((-8.0, 3.0), (8.0, 3.0), Feasibility(box, 1.0), rng, step_size=2.0, neighbourhood_radius=4.0,
                    max_iterations=1500, bounds=((-15.0, -15.0), (15.0, 15.0))) is None


@needs_reference
def test_solo_gp_planned_and_driven_to_its_goal_on_the_oracle():
    from tests.test_simulation_host import OracleAsWorld

    sc = ReferenceScenario.from_directory(SOLO)
    assert sc.rrt == tomllib.load(open(os.path.join(SOLO, "config.toml"), "rb"))["rrt"]
    planner = RRTStarPlanner.from_config(sc.rrt)
    assert (planner.step_size, planner.collision_radius, planner.neighbourhood_radius) == (5.0, 3.0, 8.0)
    sim = Simulation(sc, OracleAsWorld(sc.cfg, sc.env), np.random.default_rng(0), global_planner=planner)
    steps = sim.run(max_ticks=3000)
    m = sim.clock.missions[0]
    assert m.completed and steps < 3000, (m.state, steps)
    route = m.routes[0]
    f = Feasibility(environment_colliders(sc.env), planner.collision_radius)
    assert len(route.waypoints) >= 3 and all(f.segment(a, b, 0.5) for a, b in zip(route.waypoints, route.waypoints[1:]))
    d = sim.export()
    pos = np.asarray(d["robots"]["0"]["positions"])
    assert np.linalg.norm(pos[-1] - np.asarray(m.taskpoints[-1])) < 3.0  # arrived (finished-when: robot radius around `current`)
    assert len(d["collisions"]["environment"]) == 0, "the planned path and the Obstacle factors keep the robot off the walls"
    assert d["robots"]["0"]["mission"]["routes"][0]["waypoints"][0] == [m.taskpoints[0][0], m.taskpoints[0][1]]
