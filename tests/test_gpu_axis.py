"""The two iterate kernels (gbp_iterate_axis.cuh / gbp_iterate.cuh) against each other and the oracle.

k_iterate_axis (two lanes per variable) runs the robots whose x and y chains are decoupled, k_iterate the
rest; which kernel runs a robot is decided on the device, per launch.  Both must produce the bits of the
reference's update order, so a world iterated by both (`auto`) and a world iterated by k_iterate alone
(`general_only`) must agree bit for bit — beliefs, covariances, validity, positions, connectivity — and both
with the oracle.  The tests also pin WHICH kernel ran: a lattice away from obstacles must end up in
k_iterate_axis, a crossing swarm must leave it and come back.
"""
import numpy as np
import pytest

from magics_b200 import GbpConfig, World, scenarios
from magics_b200.sharded import LocalShards
from oracle.oracle import OracleWorld
from tests.test_gpu_parity import check
from tests.test_gpu_shards import assert_same_bits

pytestmark = pytest.mark.gpu


def _pair(sw):
    a, b = World(sw.cfg), World(sw.cfg)
    b.set_iterate_path(True)
    sw.add_to(a)
    sw.add_to(b)
    return a, b


def _same(a, b, what):
    assert_same_bits(a.read_beliefs(), b.read_beliefs(), what)
    assert np.array_equal(a.read_positions(), b.read_positions()), what
    for x, y in zip(a.read_connections(), b.read_connections()):
        assert np.array_equal(x, y), what


def test_lattice_ends_up_in_the_axis_kernel_and_matches_both_references():
    sw = scenarios.lattice(24, 15)
    a, b = _pair(sw)
    o = OracleWorld(sw.cfg)
    sw.add_to(o)
    seen = []
    for tick in range(6):
        for w in (a, b, o):
            w.step()
        _same(a, b, f"lattice tick {tick}")
        seen.append(a.read_iterate_path())
        if tick in (0, 2, 5):
            check(a, o, f"lattice tick {tick}")
    assert b.read_iterate_path() == (0, sw.n)
    assert seen[0][1] >= 0 and seen[-1] == (sw.n, 0), seen  # every robot handed to k_iterate_axis
    # half-step API and an odd schedule keep working from that state
    for w in (a, b, o):
        w.internal_factor_iteration()
        w.internal_variable_iteration()
        w.external_factor_iteration()
        w.external_variable_iteration()
        w.external_factor_iteration()
        w.external_variable_iteration()
        w.internal_factor_iteration()
        w.internal_variable_iteration()
    _same(a, b, "lattice half steps")
    check(a, o, "lattice half steps")
    assert a.read_iterate_path() == (sw.n, 0)


@pytest.mark.parametrize("speed,V", [(4.0, 11), (5.0, 12), (15.0, 21), (1.0, 5), (0.5, 3), (8.0, 15)])
def test_axis_kernel_geometries(speed, V):
    """CTA shapes: V = 11 -> 16 robots / 352 threads, 12 -> 8 / 192, 21 -> 16 / 672, 5 -> 16 / 160, 3 -> 32 / 192,
    15 -> 16 / 480."""
    cfg = GbpConfig(target_speed=speed)
    sw = scenarios.lattice(9, 7, cfg=cfg)
    assert sw.cfg.num_variables == V
    a, b = _pair(sw)
    o = OracleWorld(sw.cfg)
    sw.add_to(o)
    for tick in range(4):
        for w in (a, b, o):
            w.step()
    _same(a, b, f"V={V}")
    check(a, o, f"V={V}")
    assert a.read_iterate_path() == (sw.n, 0)


def test_crossing_circle_leaves_the_axis_kernel_and_comes_back():
    # radius 40: the horizon variables (20 m ahead) start 20 m from the centre, 7.8 m apart — outside the
    # safety distance 2.2 * (1.5 + 1.5) = 6.6 m — and close in as the robots advance 0.4 m per tick
    sw = scenarios.circle(16, circle_radius=40.0)
    a, b = _pair(sw)
    o = OracleWorld(sw.cfg)
    sw.add_to(o)
    hist = []
    for tick in range(110):
        for w in (a, b, o):
            w.step()
        hist.append(a.read_iterate_path()[0])
        if tick % 10 == 0 or tick == 109:
            _same(a, b, f"crossing tick {tick}")
            check(a, o, f"crossing tick {tick}")
    # decoupled on the way in, coupled while the InterRobot factors are active
    assert max(hist[:8]) == sw.n, hist
    assert min(hist) < sw.n // 2, hist


def test_comms_failures_and_idle_robots_in_the_axis_kernel():
    """Frozen edges (robot.rs:1851), idle robots carried over, antenna-off robots skipping the external half."""
    sw = scenarios.lattice(12, 9)
    a, b = _pair(sw)
    o = OracleWorld(sw.cfg)
    sw.add_to(o)
    rng = np.random.default_rng(5)
    for tick in range(12):
        ant = (rng.random(sw.n) > (0.3 if tick >= 3 else 0.0)).astype(np.uint8)
        idle = (rng.random(sw.n) < (0.15 if 5 <= tick < 9 else 0.0)).astype(np.uint8)
        for w in (a, b, o):
            w.set_comms(ant, idle)
            w.step()
        _same(a, b, f"comms tick {tick}")
        if tick % 3 == 2:
            check(a, o, f"comms tick {tick}")
    ax, gen = a.read_iterate_path()
    assert ax > sw.n // 2, (ax, gen)


def test_obstacles_keep_nearby_robots_in_the_general_kernel():
    sw = scenarios.complex_environment(19)
    a, b = _pair(sw)
    o = OracleWorld(sw.cfg)
    sw.add_to(o)
    for tick in range(8):
        for w in (a, b, o):
            w.step()
        _same(a, b, f"complex tick {tick}")
    check(a, o, "complex")
    # a lattice on a real SDF far from every obstacle: all in the axis kernel
    sw2 = scenarios.lattice(8, 8, pitch=0.4)
    sw2.sdf = sw.sdf
    sw2.cfg.world_width, sw2.cfg.world_height = sw.cfg.world_width, sw.cfg.world_height
    a2, b2 = _pair(sw2)
    for tick in range(4):
        a2.step()
        b2.step()
    _same(a2, b2, "lattice on the complex SDF")


def test_switching_paths_and_prior_changes_mid_run():
    sw = scenarios.lattice(10, 10)
    a, b = _pair(sw)
    o = OracleWorld(sw.cfg)
    sw.add_to(o)
    for tick in range(4):
        for w in (a, b, o):
            w.step()
    assert a.read_iterate_path() == (sw.n, 0)
    robots = np.array([3, 17, 55], np.int32)
    means = np.array([[1.0, 2.0, 0.1, 0.2], [5.0, -3.0, 0.0, 0.4], [0.0, 0.0, 0.0, 0.0]])
    for w in (a, b, o):
        w.change_prior_of_variable(4, robots, means)
        w.iterate()
    _same(a, b, "after change_prior")
    check(a, o, "after change_prior")
    a.set_iterate_path(True)   # everything through k_iterate ...
    for w in (a, b, o):
        w.step()
    assert a.read_iterate_path() == (0, sw.n)
    a.set_iterate_path(False)  # ... and back: robots re-qualify within two internal halves
    for w in (a, b, o):
        w.step()
    _same(a, b, "after switching back")
    check(a, o, "after switching back")
    assert a.read_iterate_path() == (sw.n, 0)
    # a reset (FactorGraph::reset_variables) resolves the lazily kept covariances first
    rs = np.array([0, 42], np.int32)
    rm = np.tile(np.array([0.5, 0.5, 0.0, 0.0]), (2, sw.cfg.num_variables, 1))
    for w in (a, b, o):
        w.reset_variables(rs, rm)
    _same(a, b, "after reset_variables")
    check(a, o, "after reset_variables")
    for w in (a, b, o):
        w.step()
    _same(a, b, "a tick after reset_variables")
    check(a, o, "a tick after reset_variables")


def test_sharded_lattice_uses_the_axis_kernel_and_equals_the_single_gpu_bits():
    sw = scenarios.lattice(16, 12)
    a = World(sw.cfg)
    sw.add_to(a)
    c = LocalShards(sw.cfg, 3)
    sw.add_to(c)
    for tick in range(5):
        a.step()
        c.step()
    assert_same_bits(c.read_beliefs(), a.read_beliefs(), "lattice ws=3 vs single")
    assert a.read_iterate_path() == (sw.n, 0)
    assert sum(w.read_iterate_path()[0] for w in c.shards) == sw.n


def test_fullsize_lattice_is_entirely_in_the_axis_kernel():
    sw = scenarios.lattice(400, 250)
    a = World(sw.cfg)
    sw.add_to(a)
    for _ in range(3):
        a.step()
    assert a.read_iterate_path() == (sw.n, 0)
    ba = a.read_beliefs()
    assert ba["valid"].all() and all(np.isfinite(ba[k]).all() for k in ("eta", "lam", "mean", "cov"))
    b = World(sw.cfg)
    b.set_iterate_path(True)
    sw.add_to(b)
    for _ in range(3):
        b.step()
    assert_same_bits(b.read_beliefs(), ba, "lattice-100k general_only vs auto")


def test_step_equals_its_four_calls_made_one_by_one():
    """gbp_world_step makes both prior updates in one launch (k_prior_both) and starts the next tick's neighbour
    search early; a world driven by the four calls of a tick one by one (two prior kernels, and robots removed /
    added between ticks so that a search in flight has to be thrown away) must hold the same bits."""
    sw = scenarios.circle(14, 9.0)
    a, b = World(sw.cfg), World(sw.cfg)
    o = OracleWorld(sw.cfg)
    for w in (a, b, o):
        sw.add_to(w)
    for tick in range(24):
        a.step()
        b.update_topology()
        b.update_prior_of_horizon_state()
        b.update_prior_of_current_state()
        b.iterate()
        o.step()
        if tick == 9:
            for w in (a, b, o):
                w.remove_robots([3, 7])
        if tick == 15:
            for w in (a, b, o):
                sw.add_to(w, set_sdf=False)  # a second copy of the swarm on top of the first
        if tick % 4 == 3 or tick in (10, 16):
            _same(a, b, f"tick {tick}")
            check(a, o, f"tick {tick}")


@pytest.mark.parametrize("scenario", ["circle", "junction", "lattice"])
def test_single_launch_tick_has_the_bits_of_the_launch_per_half_step_path(scenario):
    """gbp_world_set_iterate_path(w, 2): a whole iterate_gbp as one cooperative launch (k_tick_fused) for swarms that fit
    the GPU at once.  Same device functions, grid barriers where the launch boundaries were: the bits must be those
    of the default path and of the oracle, for every schedule shape, with robots added and removed in between."""
    sw = {"circle": lambda: scenarios.circle(14, 9.0), "junction": lambda: scenarios.junction_twoway(per_lane=1),
          "lattice": lambda: scenarios.lattice(12, 9)}[scenario]()
    a, c, o = World(sw.cfg), World(sw.cfg), OracleWorld(sw.cfg)
    c.set_single_launch_tick(True)
    for w in (a, c, o):
        sw.add_to(w)
    launches = []
    for tick in range(16):
        if tick == 6:
            for w in (a, c, o):
                w.set_schedule(0, 3, 5)  # centered, more external than internal halves: E E I E I E ...
        if tick == 9:
            for w in (a, c, o):
                w.set_schedule(1, 10, 10)
                w.remove_robots([2])
        if tick == 11:
            for w in (a, c, o):
                sw.slice(0, 3).add_to(w, set_sdf=False)
        l0 = (a.kernel_launches, c.kernel_launches)
        for w in (a, c, o):
            w.step()
        launches.append((a.kernel_launches - l0[0], c.kernel_launches - l0[1]))
        _same(a, c, f"{scenario} tick {tick}")
        if tick % 5 == 4 or tick == 15:
            check(c, o, f"{scenario} tick {tick}")
    assert all(lc + 8 < la for la, lc in launches[1:]), launches  # one launch instead of ~20-30 for the iterations
    assert c.read_iterate_path() == (0, c.num_robots)
