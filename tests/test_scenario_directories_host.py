"""The reference's on-disk scenario format read by the product's host side: every `config/scenarios/<name>/` directory
the reference ships (config.toml + formation.yaml + environment.yaml) goes through
`magics_b200.scenarios.ReferenceScenario.from_directory`, spawns its formations and — on the CPU oracle — runs its first
ticks with the SDF generated from its own environment.  The three BASELINE scenarios must parse to exactly the committed
tests/golden/scenarios.json (which tests/golden/make_golden.py writes with the same reader).  The reference tree exists
only in the build container; on the GPU box the `-m gpu` twin (tests/test_gpu_reference_scenarios.py) reads the golden file."""
import json
import os

import numpy as np
import pytest

from magics_b200 import scenarios
from magics_b200.scenarios import ReferenceScenario, read_scenario_directory

REF = "/root/reference/config/scenarios"
pytestmark = pytest.mark.skipif(not os.path.isdir(REF), reason="the reference tree exists only in the build container")
NAMES = sorted(os.listdir(REF)) if os.path.isdir(REF) else []
# variables per robot the scenario's own target speed and planning horizon give (utils.rs:35-75); the engine keeps a
# robot's variables inside one warp and refuses V > 32 - two experiments of the thesis sit above
ENGINE_MAX_V = 32


def test_the_three_baseline_scenarios_parse_to_the_committed_golden(golden_dir):
    golden = json.load(open(os.path.join(golden_dir, "scenarios.json")))
    assert set(golden) == {"Circle Experiment", "Structured Junction Twoway", "Collaborative Complex"}
    for name, want in golden.items():
        got = json.loads(json.dumps(read_scenario_directory(os.path.join(REF, name))))
        assert got == want, name


@pytest.mark.parametrize("name", NAMES)
def test_every_shipped_scenario_loads_spawns_and_steps_on_the_oracle(name):
    from oracle import oracle
    from oracle.oracle import OracleWorld

    sc = ReferenceScenario.from_directory(os.path.join(REF, name))
    assert sc.cfg.num_variables == len(sc.timesteps) >= 3 and sc.world_w > 0 and sc.world_h > 0
    assert len(sc.formations) >= 1
    ticks = 12
    events = sc.spawn_events(int(30 * sc.hz))  # half a minute of the spawner clock
    rng = np.random.default_rng(0)
    if all(f.robots == 0 for f in sc.formations):  # Obstacle Shapes Showcase: an environment without robots
        assert all(sc.spawn(k, rng) is None for _, k in events)
        img = oracle.env_to_sdf_image(sc.env)
        assert img.shape[2] == 3 and img.min() < 128 < img.max()
        return
    assert events and events == sorted(events)
    o = OracleWorld(sc.cfg, threads=4)
    o.set_sdf(oracle.env_to_sdf_image(sc.env))
    first_tick = events[0][0]
    spawned = 0
    for tick in range(first_tick, first_tick + ticks):
        for _, k in [e for e in events if e[0] == tick]:
            sw = sc.spawn(k, rng)
            if sw is None:  # "failed to spawn formation, skipping" (random placement ran out of attempts)
                continue
            assert sw.n == sc.formations[k].robots and sw.init_means.shape == (sw.n, sc.cfg.num_variables, 4)
            # every robot starts inside the world and has at least one waypoint beyond its start
            assert (np.abs(sw.positions[:, 0]) <= sc.world_w).all() and (np.abs(sw.positions[:, 1]) <= sc.world_h).all()
            assert (np.diff(sw.wp_offsets) >= 2).all()
            sw.add_to(o, set_sdf=False)
            spawned += sw.n
        if o.num_robots:
            if sc.reached_when is not None:
                o.reached_waypoint(sc.reached_when, sc.finished_when)
            o.step()
    assert spawned >= 1 and o.num_robots == spawned
    b = o.read_beliefs()
    assert np.isfinite(b["mean"]).all(), name
    # the robots moved: the current-state variable has left its spawn point by up to 12 ticks of target speed
    moved = np.linalg.norm(o.read_positions() - b["mean"][:, 0, :2].astype(np.float32), axis=1)
    assert (moved <= 1e-3).all()  # Transform follows variable 0 (robot.rs:2286-2338)
    if sc.cfg.num_variables > ENGINE_MAX_V:
        assert name in ("Communications Failure Experiment", "Varying Network Connectivity Experiment")


def test_communications_failure_rate_is_read_and_drawn_per_robot_and_tick():
    sc = ReferenceScenario.from_directory(os.path.join(REF, "Communications Failure Experiment"))
    assert 0.0 <= sc.failure_rate <= 1.0
    raw = read_scenario_directory(os.path.join(REF, "Communications Failure Experiment"))
    assert sc.failure_rate == float(raw["robot"]["communication"]["failure-rate"])
    sc.failure_rate = 0.3
    rng = np.random.default_rng(1)
    draws = np.stack([sc.draw_antennas(2000, rng) for _ in range(10)])
    assert draws.dtype == np.uint8 and abs(1.0 - draws.mean() - 0.3) < 0.02 and (draws[0] != draws[1]).any()
    sc.failure_rate = 0.0
    assert sc.draw_antennas(5, rng).tolist() == [1] * 5
