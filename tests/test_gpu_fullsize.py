"""BASELINE.json's full sizes (config 4: 100 000-robot rings, config 5: 1 000 000-robot lattice), where the
oracle would take hours: size-independent properties instead —
  * connectivity is symmetric and has the degrees the geometry dictates,
  * robot_number is exactly the reference's creation order (robots in id order -> neighbours ascending ->
    i = 1..V-1, robot.rs:1500-1541): 1 + (V-1) * (position of the directed pair in CSR order),
  * two independent runs give identical bits (no atomics / no order dependence on the float path),
  * the swarm partitioned over several shards gives the bits of the single-GPU engine,
  * every belief is finite and valid, and the plan actually moved.
Edge cases at the other end (no robots, one robot, V = 2, coincident robots) run against the oracle."""
import numpy as np
import pytest

from magics_b200 import GbpConfig, World, scenarios
from magics_b200.sharded import LocalShards
from oracle.oracle import OracleWorld
from tests.test_gpu_parity import check
from tests.test_gpu_shards import assert_same_bits

pytestmark = pytest.mark.gpu


def _run(sw, ticks, world=None):
    g = world or World(sw.cfg)
    sw.add_to(g)
    for _ in range(ticks):
        g.step()
    return g


def _assert_symmetric(off, nb):
    n = off.shape[0] - 1
    src = np.repeat(np.arange(n, dtype=np.int64), np.diff(off))
    fwd = src * n + nb
    bwd = nb.astype(np.int64) * n + src
    assert np.array_equal(np.sort(fwd), np.sort(bwd)), "connectivity is not symmetric"
    assert np.all(nb != src), "self edge"
    # neighbours ascending inside every row (BTreeSet order)
    inner = np.ones(nb.shape[0], bool)
    inner[off[1:-1][off[1:-1] < nb.shape[0]]] = False
    assert np.all((np.diff(nb) > 0) | ~inner[1:]), "neighbour lists are not sorted"


def test_rings_100k_properties_determinism_and_shards():
    sw = scenarios.rings(100_000)
    V = sw.cfg.num_variables
    a = _run(sw, 2)
    off, nb, rn = a.read_connections()
    _assert_symmetric(off, nb)
    deg = np.diff(off)
    assert deg.min() >= 2 and deg.max() <= 6 and abs(deg.mean() - 4.0) < 0.05
    assert np.array_equal(rn, 1 + (V - 1) * np.arange(nb.shape[0], dtype=np.int64)), "robot_number order"
    ba = a.read_beliefs()
    assert ba["valid"].all() and all(np.isfinite(ba[k]).all() for k in ("eta", "lam", "mean", "cov"))
    assert np.abs(ba["mean"][:, 1:-1] - sw.init_means[:, 1:-1]).max() > 1e-4, "beliefs never moved"
    b = _run(sw, 2)
    assert_same_bits(b.read_beliefs(), ba, "rings-100000 run 2 vs run 1")
    b.close()
    c = _run(sw, 2, LocalShards(sw.cfg, 4))
    assert_same_bits(c.read_beliefs(), ba, "rings-100000 ws=4 vs single")
    for x, y in zip(c.read_connections(), (off, nb, rn)):
        assert np.array_equal(x, y)


def test_lattice_1m_properties_and_two_shards():
    nx = ny = 1000
    sw = scenarios.lattice(nx, ny)
    V = sw.cfg.num_variables
    a = _run(sw, 1)
    off, nb, rn = a.read_connections()
    # 8-neighbourhood at pitch 12 / comms 20: horizontal + vertical + both diagonals, directed
    expect = 2 * ((nx - 1) * ny + nx * (ny - 1) + 2 * (nx - 1) * (ny - 1))
    assert nb.shape[0] == expect
    deg = np.diff(off).reshape(ny, nx)
    assert (deg[1:-1, 1:-1] == 8).all() and deg[0, 0] == 3 and deg[0, 1] == 5 and deg[-1, -1] == 3
    _assert_symmetric(off, nb)
    assert np.array_equal(rn, 1 + (V - 1) * np.arange(nb.shape[0], dtype=np.int64)), "robot_number order"
    mean_a = a.read_beliefs(eta=False, lam=False, cov=False)
    assert mean_a["valid"].all() and np.isfinite(mean_a["mean"]).all()
    a.close()
    c = _run(sw, 1, LocalShards(sw.cfg, 2))
    mean_c = c.read_beliefs(eta=False, lam=False, cov=False)
    assert np.array_equal(mean_c["mean"], mean_a["mean"]) and np.array_equal(mean_c["valid"], mean_a["valid"])
    o2, n2, r2 = c.read_connections()
    assert np.array_equal(o2, off) and np.array_equal(n2, nb) and np.array_equal(r2, rn)
    assert [w.num_ghosts for w in c.shards] == [nx, nx]


def test_empty_world_and_single_robot():
    cfg = scenarios.circle(1).cfg
    g = World(cfg)
    g.step()  # no robots: every system is a no-op
    g.update_topology()
    g.iterate()
    assert g.num_robots == 0 and g.read_connections()[0].tolist() == [0]
    sw = scenarios.circle(1)
    o = OracleWorld(sw.cfg)
    sw.add_to(g)
    sw.add_to(o)
    for _ in range(5):
        g.step()
        o.step()
    check(g, o, "one robot")


def test_two_variables_per_robot_minimum():
    """V = 2 (lookahead horizon 1): no Obstacle/Tracking factors exist, 16 robots share a warp."""
    from magics_b200 import get_variable_timesteps

    ts = get_variable_timesteps(1, 1)
    assert ts.tolist() == [0, 1]
    cfg = GbpConfig(target_speed=1.0, world_width=100.0, world_height=100.0)
    sw = scenarios.circle(40, circle_radius=30.0, cfg=cfg, planning_horizon=1.0, lookahead_multiple=1)
    assert sw.cfg.num_variables == 2
    g, o = World(sw.cfg), OracleWorld(sw.cfg)
    sw.add_to(g)
    sw.add_to(o)
    for tick in range(6):
        g.step()
        o.step()
        check(g, o, f"V=2 tick {tick}")


def test_coincident_robots_and_exact_range_boundary():
    """Robots on the same spot (distance 0) and pairs exactly at the comms radius (kept: the reference
    tests `radius < distance`, robot.rs:1373)."""
    sw = scenarios.circle(6, circle_radius=10.0)
    sw.positions[1] = sw.positions[0]
    sw.init_means[1] = sw.init_means[0]
    sw.positions[2] = sw.positions[0] + np.array([20.0, 0.0], np.float32)  # exactly the comms radius
    sw.init_means[2, :, 0] = sw.init_means[0, :, 0] + 20.0
    sw.init_means[2, :, 1] = sw.init_means[0, :, 1]
    g, o = World(sw.cfg), OracleWorld(sw.cfg)
    sw.add_to(g)
    sw.add_to(o)
    g.update_topology()
    o.update_topology()
    off, nb, _ = g.read_connections()
    assert 2 in nb[off[0]:off[1]] and 1 in nb[off[0]:off[1]]
    for tick in range(8):
        g.step()
        o.step()
        check(g, o, f"coincident tick {tick}")


def test_long_run_through_the_crossing_stays_identical():
    """150 ticks = 1500 robot-GBP-iterations per robot: 16 robots cross the centre of a 25 m circle, with
    InterRobot factors switching on and off, edges being created and deleted and radios failing now and
    then.  No drift between engine and oracle is tolerated anywhere along the way (1e-9, measured 0.0)."""
    sw = scenarios.circle(16, circle_radius=25.0)
    g, o = World(sw.cfg), OracleWorld(sw.cfg, threads=8)
    sw.add_to(g)
    sw.add_to(o)
    rng = np.random.default_rng(17)
    closest = np.inf
    for tick in range(150):
        ant = (rng.uniform(size=16) > 0.05).astype(np.uint8)
        for w in (g, o):
            w.set_comms(ant, None)
            w.step()
        if tick % 10 == 9:
            check(g, o, f"long run tick {tick}")
            p = o.read_positions().astype(np.float64)
            d = np.linalg.norm(p[:, None] - p[None], axis=-1) + np.eye(16) * 1e9
            closest = min(closest, d.min())
    assert closest < 6.0, "the robots never got close: the test did not exercise the InterRobot factors"
    assert np.linalg.norm(o.read_positions()[0] - sw.positions[0]) > 35.0  # it crossed
