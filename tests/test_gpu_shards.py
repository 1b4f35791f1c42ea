"""Parity of the SHARDED engine: one swarm partitioned over several shards must give
the reference's results bit for bit in connectivity / robot_number and within 1e-9 in beliefs —
and, stronger, exactly the bits of the unsharded engine, since every shard runs the same arithmetic
in the same order on the same operands.

These tests drive `world_size` shards inside one process on ONE GPU (`LocalShards`,
gbp_world_create_local_shards): the whole multi-GPU code path — ghost slots, per-peer send lists,
cross-shard robot_number exchange, per-sub-step halo pack/exchange/unpack — with device-to-device
copies as the transport.  tests/test_gpu_nccl.py runs the same path over NCCL when the box has two GPUs."""
import numpy as np
import pytest

from magics_b200 import SCHEDULE_CENTERED, SCHEDULE_INTERLEAVE_EVENLY, GbpConfig, World, scenarios
from magics_b200.sharded import LocalShards
from oracle.oracle import OracleWorld
from tests.parity import assert_beliefs_match
from tests.test_gpu_parity import check

pytestmark = pytest.mark.gpu


def make_pair(sw, ws, bounds=None):
    g = LocalShards(sw.cfg, ws, bounds=bounds)
    o = OracleWorld(sw.cfg)
    sw.add_to(g)
    sw.add_to(o)
    return g, o


def assert_same_bits(a: dict, b: dict, what: str):
    for k in ("eta", "lam", "mean", "cov", "valid"):
        assert np.array_equal(a[k], b[k], equal_nan=True), f"{what}: {k} differs between sharded and single-GPU engine"


@pytest.mark.parametrize("ws", [2, 3, 4])
def test_lattice_slabs_connectivity_robot_numbers_and_beliefs(ws):
    """Config 5 shape: row-major lattice cut into `ws` horizontal slabs; every slab boundary is
    crossed by 3 InterRobot factors per border robot."""
    sw = scenarios.lattice(20, 15)
    g, o = make_pair(sw, ws)
    g.update_topology()
    o.update_topology()
    og, ng, rg = g.read_connections()
    oo, no, ro = o.read_connections()
    assert np.array_equal(og, oo) and np.array_equal(ng, no), "connectivity differs"
    assert np.array_equal(rg, ro), "robot_number differs across shards"
    assert sum(w.num_ghosts for w in g.shards) > 0
    for tick in range(3):
        g.step()
        o.step()
    check(g, o, f"lattice ws={ws}")


def test_circle_30_crossing_four_shards():
    """Config 1 on 4 shards: ghosts and send lists change every few ticks as the robots converge."""
    sw = scenarios.circle(30)
    g, o = make_pair(sw, 4)
    ghosts = set()
    for tick in range(25):
        g.step()
        o.step()
        ghosts.add(tuple(w.num_ghosts for w in g.shards))
        if tick % 4 == 0 or tick == 24:
            check(g, o, f"circle ws=4 tick {tick}")
    assert len(ghosts) > 1, "ghost sets never changed: the test does not exercise re-sharding"


def test_topology_create_delete_uneven_shards_with_an_empty_one():
    sw = scenarios.circle(10, circle_radius=16.0)
    sw.cfg.comms_radius = 9.0
    g, o = make_pair(sw, 4, bounds=[0, 3, 3, 4, 10])  # shard 1 owns nothing, shard 2 one robot
    sizes = []
    for tick in range(60):
        g.step()
        o.step()
        if tick % 6 == 0:
            check(g, o, f"crossing ws=4 tick {tick}")
            sizes.append(g.read_connections()[1].size)
    check(g, o, "crossing end")
    assert max(sizes) > sizes[0] and sizes[-1] < max(sizes), sizes


def test_comms_failure_and_idle_masks_across_shards():
    sw = scenarios.circle(10, circle_radius=16.0)
    sw.cfg.comms_radius = 11.0
    g, o = make_pair(sw, 3)
    rng = np.random.default_rng(11)
    for tick in range(40):
        ant = (rng.uniform(size=10) > 0.25).astype(np.uint8)
        idle = (rng.uniform(size=10) > 0.85).astype(np.uint8) if 5 <= tick < 15 else None
        for w in (g, o):
            w.set_comms(ant, idle)
            w.step()
        if tick % 5 == 4:
            check(g, o, f"idle+comms ws=3 tick {tick}")


def test_junction_twoway_all_factor_kinds_two_shards():
    sw = scenarios.junction_twoway(per_lane=2)
    g, o = make_pair(sw, 2)
    for tick in range(16):
        g.step()
        o.step()
        if tick % 3 == 0 or tick == 15:
            check(g, o, f"junction ws=2 tick {tick}")


def test_change_prior_setters_and_half_steps_across_shards():
    sw = scenarios.circle(6, circle_radius=10.0)
    g, o = make_pair(sw, 3)
    for w in (g, o):
        w.step()
        w.change_prior_of_variable(3, [0, 4], np.array([[1.0, 2.0, 0.5, 0.25], [-3.0, 1.0, 0.0, 0.1]]))
        w.set_safety_distance_multiplier(3.0)
        w.step()
        w.change_factor_enabled(2, 0)
        w.set_schedule(SCHEDULE_CENTERED, 6, 3)
        w.step()
    check(g, o, "setters ws=3")
    for k in range(3):
        for w in (g, o):
            w.internal_factor_iteration()
            w.internal_variable_iteration()
            w.external_factor_iteration()
            w.external_variable_iteration()
        check(g, o, f"half steps ws=3 {k}")


@pytest.mark.parametrize("kind,internal,external", [(SCHEDULE_INTERLEAVE_EVENLY, 3, 8), (SCHEDULE_CENTERED, 10, 5),
                                                    (SCHEDULE_INTERLEAVE_EVENLY, 0, 3)])
def test_schedules_with_consecutive_external_halves(kind, internal, external):
    sw = scenarios.circle(8, circle_radius=12.0)
    sw.cfg.schedule_kind, sw.cfg.iterations_internal, sw.cfg.iterations_external = kind, internal, external
    g, o = make_pair(sw, 2)
    for tick in range(5):
        g.step()
        o.step()
        check(g, o, f"schedule {kind} {internal}/{external} ws=2 tick {tick}")


def test_sharded_equals_single_gpu_bitwise_rings_20000():
    """Beyond oracle sizes: the sharded and the single-GPU engine must agree in every bit."""
    sw = scenarios.rings(20000)
    one = World(sw.cfg)
    sw.add_to(one)
    many = LocalShards(sw.cfg, 4)
    sw.add_to(many)
    for tick in range(3):
        one.step()
        many.step()
    assert_same_bits(many.read_beliefs(), one.read_beliefs(), "rings-20000 ws=4")
    a, b = many.read_connections(), one.read_connections()
    for x, y in zip(a, b):
        assert np.array_equal(x, y)
    assert sum(w.num_ghosts for w in many.shards) >= 6  # every cut of a ring has ghosts on both sides


def test_sharded_equals_single_gpu_bitwise_lattice_moving():
    """A lattice whose rows drift at different speeds: cross-slab edges are created and deleted."""
    sw = scenarios.lattice(40, 24, pitch=14.0)
    # even rows head +x, odd rows -x: neighbours in adjacent rows change as the rows shear
    wp = sw.wp_xy.reshape(-1, 2, 2).copy()
    rows = (np.arange(sw.n) // 40) % 2 == 1
    wp[rows, 1, 0] = wp[rows, 0, 0] - 500.0
    sw.wp_xy = wp.reshape(-1, 2)
    sw.init_means = scenarios.initial_means(wp[:, 0], wp[:, 1], sw.timesteps, sw.cfg.target_speed, 5.0)
    one = World(sw.cfg)
    sw.add_to(one)
    many = LocalShards(sw.cfg, 3, bounds=[0, 40 * 7, 40 * 16, 40 * 24])
    sw.add_to(many)
    edges = []
    for tick in range(30):
        one.step()
        many.step()
        if tick % 10 == 9:
            assert_same_bits(many.read_beliefs(), one.read_beliefs(), f"shear tick {tick}")
            a, b = many.read_connections(), one.read_connections()
            for x, y in zip(a, b):
                assert np.array_equal(x, y), f"shear tick {tick}: connectivity / robot_number"
            edges.append(a[1].size)
    assert np.array_equal(many.read_positions(), one.read_positions())


def test_robots_spawned_after_commit_join_the_last_shard():
    """FormationSpawner repeats (spawner.rs:186-323): robots that appear while the simulation runs take the ids above
    every existing one, i.e. they join the last shard whatever their position, and the group publishes the new total
    (gbp_world_commit_shards again).  Connectivity, robot_number and beliefs must stay those of the oracle and the
    bits those of a single-GPU world that received the same robots at the same tick."""
    sw = scenarios.circle(12, 14.0)
    late = scenarios.circle(9, 11.0)  # a second wave, spawned inside the first one's circle
    g, o = make_pair(sw, 3)
    one = World(sw.cfg)
    sw.add_to(one)
    worlds = (g, one, o)
    for tick in range(30):
        for w in worlds:
            w.step()
        if tick == 7:
            for w in worlds:
                late.add_to(w, set_sdf=False)
        if tick == 15:
            for w in worlds:
                late.add_to(w, set_sdf=False)
        if tick in (7, 8, 12, 16, 22, 29):
            check(g, o, f"late spawn tick {tick}")
            assert_same_bits(g.read_beliefs(), one.read_beliefs(), f"late spawn tick {tick}")
            assert np.array_equal(g.read_positions(), one.read_positions())
    assert g.num_robots == sw.n + 2 * late.n and g.shards[-1].num_robots == sw.n // 3 + 2 * late.n
    with pytest.raises(RuntimeError, match="last shard"):
        late.slice(0, 2).add_to(g.shards[0], set_sdf=False)
