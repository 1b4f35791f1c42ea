"""BASELINE configs 1-3 on the reference's OWN inputs (tests/golden/scenarios.json: formation.yaml, environment.yaml
and config.toml scalars of `Circle Experiment`, `Structured Junction Twoway`, `Collaborative Complex`):
robots are spawned by the FormationSpawner clock through Formation::as_positions (magics_b200/formation.py), the SDF
is generated on the device from the environment (gbp_world_set_sdf_from_environment) for the engine and by the CPU
restatement for the oracle, waypoints are reached with the scenario's criteria and finished robots despawn.
Engine == oracle at every check."""
import numpy as np
import pytest

from magics_b200 import World, scenarios
from oracle import oracle
from oracle.oracle import OracleWorld
from tests.test_gpu_parity import check

pytestmark = pytest.mark.gpu


def _run(name, ticks, every, seed=0):
    sc = scenarios.ReferenceScenario(name)
    g, o = World(sc.cfg), OracleWorld(sc.cfg)
    g.set_sdf_from_environment(sc.env)
    o.set_sdf(oracle.env_to_sdf_image(sc.env))
    rng = np.random.default_rng(seed)
    events = sc.spawn_events(ticks)
    gone = np.zeros(0, bool)
    nwp = np.zeros(0, np.int64)  # route points per robot: mission.is_completed() <=> next waypoint index == nwp
    spawned = despawned = 0
    max_edges = 0
    for tick in range(ticks):
        for _, k in [e for e in events if e[0] == tick]:
            sw = sc.spawn(k, rng)
            assert sw is not None
            sw.add_to(g, set_sdf=False)
            sw.add_to(o, set_sdf=False)
            gone = np.concatenate([gone, np.zeros(sw.n, bool)])
            nwp = np.concatenate([nwp, np.diff(sw.wp_offsets)])
            spawned += sw.n
        if g.num_robots == 0:
            continue
        rg = g.reached_waypoint(sc.reached_when, sc.finished_when)
        ro = o.reached_waypoint(sc.reached_when, sc.finished_when)
        assert np.array_equal(rg, ro), f"{name} tick {tick}: different robots reached a waypoint"
        if sc.despawn:
            done = (g.read_waypoint_index() >= nwp) & ~gone
            if done.any():
                ids = np.flatnonzero(done).astype(np.int32)
                g.remove_robots(ids)
                o.remove_robots(ids)
                gone |= done
                despawned += int(done.sum())
        g.step()
        o.step()
        max_edges = max(max_edges, int(g.read_connections()[0][-1]))
        if tick % every == 0 or tick == ticks - 1:
            check(g, o, f"{name} tick {tick}")
    return sc, g, o, spawned, despawned, max_edges


def test_circle_experiment_inputs():
    """Config 1 with the scenario's own scalars: 30 robots, radii drawn in [2, 3] (no uniform Dynamic-factor table),
    V = 21, 50 internal / 10 external iterations, comms radius 50."""
    sc, g, o, spawned, _, edges = _run("Circle Experiment", 26, 5)
    assert spawned == 30 and g.cfg.num_variables == 21
    assert edges > 0


def test_structured_junction_twoway_inputs():
    """Config 2: all four factor kinds, tracking along the three-point routes through the junction, four spawn waves."""
    sc, g, o, spawned, _, edges = _run("Structured Junction Twoway", 70, 10)
    assert spawned == 16 and edges > 0
    rec, pos, val = g.read_tracking()
    assert (val != 0).any()


def test_collaborative_complex_inputs():
    """Config 3: the 10 x 7 tile maze, SDF 2000 x 1400 generated on the device, Obstacle factors on."""
    sc, g, o, spawned, _, edges = _run("Collaborative Complex", 45, 9)
    assert spawned >= 12
    b = g.read_beliefs()
    assert np.isfinite(b["mean"]).all()
