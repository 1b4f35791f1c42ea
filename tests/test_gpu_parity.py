"""Parity of the CUDA engine (through the C ABI) against the CPU oracle on the
same seeded inputs — the `-m gpu` tests proper.  Bit-exact for graph indexing
(connectivity, creation order / robot_number, SDF pixel indices); beliefs and
means within 1e-9 relative (tests/parity.py states the metric)."""
import numpy as np
import pytest

from magics_b200 import (SCHEDULE_CENTERED, SCHEDULE_HALF_BEGINNING_HALF_END, SCHEDULE_INTERLEAVE_EVENLY,
                         SCHEDULE_LATE_AS_POSSIBLE, SCHEDULE_SOON_AS_POSSIBLE, GbpConfig, World, scenarios)
from oracle.oracle import OracleWorld
from tests.parity import assert_beliefs_match

pytestmark = pytest.mark.gpu


def make_pair(sw):
    g = World(sw.cfg)
    o = OracleWorld(sw.cfg)
    sw.add_to(g)
    sw.add_to(o)
    return g, o


def check(g, o, what, connectivity=True):
    errs = assert_beliefs_match(g.read_beliefs(), o.read_beliefs(), what=what)
    assert np.array_equal(g.read_positions(), o.read_positions()), what + ": f32 positions differ"
    if connectivity:
        og, ng, rg = g.read_connections()
        oo, no, ro = o.read_connections()
        assert np.array_equal(og, oo) and np.array_equal(ng, no), what + ": connectivity differs"
        assert np.array_equal(rg, ro), what + ": robot_number (creation order) differs"
    assert np.array_equal(g.node_counts(), o.node_counts()), what
    return errs


def test_single_robot_dynamics_only_every_half_step():
    sw = scenarios.circle(1)
    sw.cfg.enable_obstacle = 0
    g, o = make_pair(sw)
    check(g, o, "initial state")
    for k in range(12):
        for w in (g, o):
            w.internal_factor_iteration()
            w.internal_variable_iteration()
        check(g, o, f"internal half {k}")
    for w in (g, o):
        w.external_factor_iteration()
        w.external_variable_iteration()
    check(g, o, "external half")


def test_circle_30_ticks_default_config():
    """BASELINE config 1: circle scenario, default scalars, 10/10 interleave-evenly."""
    sw = scenarios.circle(30)
    g, o = make_pair(sw)
    for tick in range(25):
        g.step()
        o.step()
        if tick % 4 == 0 or tick == 24:
            check(g, o, f"circle tick {tick}")
    off, nb, rn = g.read_connections()
    assert nb.size > 60  # robots converged: connectivity grew beyond the initial K=2


def test_circle_experiment_scalars_50_10():
    """`Circle Experiment` scenario scalars: 50 internal / 10 external, V=21, comms 50."""
    cfg = GbpConfig(sigma_factor_dynamics=1.0, sigma_factor_interrobot=0.005, sigma_factor_obstacle=0.005,
                    target_speed=15.0, comms_radius=50.0, iterations_internal=50, iterations_external=10,
                    world_width=100.0, world_height=100.0)
    sw = scenarios.circle(12, robot_radius=2.5, cfg=cfg)
    assert sw.cfg.num_variables == 21
    g, o = make_pair(sw)
    for tick in range(6):
        g.step()
        o.step()
        check(g, o, f"50/10 tick {tick}")


@pytest.mark.parametrize("kind,internal,external", [
    (SCHEDULE_CENTERED, 4, 6), (SCHEDULE_CENTERED, 10, 5), (SCHEDULE_INTERLEAVE_EVENLY, 6, 7),
    (SCHEDULE_INTERLEAVE_EVENLY, 3, 8), (SCHEDULE_SOON_AS_POSSIBLE, 5, 2), (SCHEDULE_LATE_AS_POSSIBLE, 2, 5),
    (SCHEDULE_HALF_BEGINNING_HALF_END, 5, 3), (SCHEDULE_INTERLEAVE_EVENLY, 0, 3), (SCHEDULE_INTERLEAVE_EVENLY, 3, 0),
])
def test_all_schedules(kind, internal, external):
    sw = scenarios.circle(8, circle_radius=12.0)
    sw.cfg.schedule_kind, sw.cfg.iterations_internal, sw.cfg.iterations_external = kind, internal, external
    g, o = make_pair(sw)
    for tick in range(5):
        g.step()
        o.step()
        check(g, o, f"schedule {kind} {internal}/{external} tick {tick}")


def test_open_half_iteration_pair_refuses_reads_and_changes():
    """internal_factor_iteration alone only opens a pair (the factor half runs fused with the variable half): until it
    is closed the engine refuses to show or change state the reference would already have advanced."""
    sw = scenarios.circle(4, circle_radius=6.0)
    g, o = make_pair(sw)
    g.internal_factor_iteration()
    for call in (g.read_beliefs, lambda: g.set_comms(np.ones(4, np.uint8), None), g.update_topology, g.iterate,
                 g.external_factor_iteration, lambda: g.remove_robots([0])):
        with pytest.raises(RuntimeError, match="half-iteration|order|pair"):
            call()
    g.internal_variable_iteration()
    o.internal_factor_iteration()
    o.internal_variable_iteration()
    check(g, o, "pair closed")


def test_junction_twoway_all_factor_kinds():
    """BASELINE config 2: tracking + obstacle + interrobot + dynamic, V=12."""
    sw = scenarios.junction_twoway(per_lane=2)
    assert sw.cfg.num_variables == 12 and sw.cfg.enable_tracking
    g, o = make_pair(sw)
    for tick in range(16):
        g.step()
        o.step()
        if tick % 3 == 0 or tick == 15:
            check(g, o, f"junction tick {tick}")
    # Tracking factor state (tracking.rs:62-90): record counter, LastMeasurement.pos (f32) and .value, bit for bit
    rec, pos, val = g.read_tracking()
    moved = 0
    for r in range(sw.n):
        for i in range(1, sw.cfg.num_variables - 1):
            t = o.read_tracking(r, i)
            assert t is not None
            assert int(rec[r, i]) == t[0], (r, i, rec[r, i], t[0])
            assert pos[r, i].tobytes() == np.asarray(t[1], np.float32).tobytes(), (r, i, pos[r, i], t[1])
            assert np.float64(val[r, i]).tobytes() == np.float64(t[2]).tobytes(), (r, i, val[r, i], t[2])
            moved += int(t[2] != 0.0)
    assert moved > 0  # the factors have run (iteration_count.factor >= 10) and measured something


def test_complex_environment_obstacle_factors():
    """BASELINE config 3: 2000x1400 SDF, obstacle factors pushing on the plan."""
    sw = scenarios.complex_environment(19)
    g, o = make_pair(sw)
    for tick in range(12):
        g.step()
        o.step()
    check(g, o, "complex")
    b = g.read_beliefs()
    assert np.abs(b["mean"][:, :, 1] - sw.init_means[:, :, 1]).max() > 1e-3


def test_sdf_lookup_bit_exact_pixels():
    sw = scenarios.complex_environment(2)
    g, o = make_pair(sw)
    rng = np.random.default_rng(0)
    xy = np.concatenate([
        rng.uniform([-130.0, -90.0], [130.0, 90.0], size=(20000, 2)),          # inside + just outside
        np.array([[0.0, 0.0], [-125.0, 87.5], [125.0, -87.5], [124.99999999, -87.49999999], [1e30, 0.0],
                  [-1e30, 0.0], [np.nan, 0.0], [0.0, np.inf], [-125.0 - 1e-13, 0.0]]),
        (np.arange(-1000, 1001)[:, None] * np.array([[0.125, 0.0]])),           # exact pixel boundaries
    ])
    pg = g.sdf_lookup(xy)
    po = o.sdf_lookup(xy)
    for a, b, name in zip(pg, po, ("px", "py", "value")):
        assert np.array_equal(a, b), name


def test_lattice_connectivity_and_robot_numbers_bit_exact():
    sw = scenarios.lattice(20, 15)
    g, o = make_pair(sw)
    g.update_topology()
    o.update_topology()
    og, ng, rg = g.read_connections()
    oo, no, ro = o.read_connections()
    assert np.array_equal(og, oo) and np.array_equal(ng, no) and np.array_equal(rg, ro)
    deg = np.diff(og)
    assert deg.max() == 8 and deg.min() == 3
    for tick in range(3):
        g.step()
        o.step()
    check(g, o, "lattice")


def test_topology_changes_create_and_delete():
    """Robots crossing: edges appear and disappear; surviving edges keep state."""
    sw = scenarios.circle(10, circle_radius=16.0)
    sw.cfg.comms_radius = 9.0
    g, o = make_pair(sw)
    sizes = []
    for tick in range(60):
        g.step()
        o.step()
        if tick % 6 == 0:
            check(g, o, f"crossing tick {tick}")
            sizes.append(g.read_connections()[1].size)
    check(g, o, "crossing end")
    assert max(sizes) > sizes[0] and sizes[-1] < max(sizes), sizes


def test_change_prior_and_setters():
    sw = scenarios.circle(6, circle_radius=10.0)
    g, o = make_pair(sw)
    for w in (g, o):
        w.step()
        w.change_prior_of_variable(3, [0, 4], np.array([[1.0, 2.0, 0.5, 0.25], [-3.0, 1.0, 0.0, 0.1]]))
        w.set_safety_distance_multiplier(3.0)
        w.step()
        w.change_factor_enabled(2, 0)
        w.set_schedule(SCHEDULE_CENTERED, 6, 3)
        w.step()
    check(g, o, "setters")


def test_comms_failure_and_idle_masks():
    sw = scenarios.circle(8, circle_radius=12.0)
    g, o = make_pair(sw)
    rng = np.random.default_rng(3)
    for w in (g, o):
        w.step()
    for tick in range(6):
        ant = (rng.uniform(size=8) > 0.3).astype(np.uint8)
        for w in (g, o):
            w.set_comms(ant, None)
            w.step()
    check(g, o, "comms failure")


def test_idle_robots_and_comms_failure_with_topology_changes():
    """mission.state.idle() robots neither iterate nor receive (robot.rs:1791,1806,1822);
    antenna failures while robots cross (edges created while a radio is off)."""
    sw = scenarios.circle(10, circle_radius=16.0)
    sw.cfg.comms_radius = 11.0
    g, o = make_pair(sw)
    rng = np.random.default_rng(11)
    for tick in range(40):
        ant = (rng.uniform(size=10) > 0.25).astype(np.uint8)
        idle = (rng.uniform(size=10) > 0.85).astype(np.uint8) if 5 <= tick < 15 else None
        for w in (g, o):
            w.set_comms(ant, idle)
            w.step()
        if tick % 5 == 4:
            check(g, o, f"idle+comms tick {tick}")


def test_robots_added_later_join_the_graph():
    sw = scenarios.circle(8, circle_radius=12.0)
    a, b = sw.slice(0, 5), sw.slice(5, 8)
    g, o = World(sw.cfg), OracleWorld(sw.cfg)
    for w in (g, o):
        a.add_to(w)
        w.step()
        w.step()
        b.add_to(w, set_sdf=False)
        for _ in range(3):
            w.step()
    check(g, o, "late spawn")


def test_rings_2000_full_iteration_parity():
    """Config 4 shape at a size the oracle finishes in seconds."""
    sw = scenarios.rings(2000)
    g, o = make_pair(sw)
    o.set_threads(8)
    for tick in range(3):
        g.step()
        o.step()
    errs = check(g, o, "rings-2000")
    assert np.diff(g.read_connections()[0]).mean() > 3.5


def test_async_readback_overlaps_the_next_tick_without_tearing():
    """gbp_world_read_beliefs_async: the means copied out while the next tick runs are the means of the
    tick they were requested after, bit for bit."""
    from magics_b200 import pinned_empty

    sw = scenarios.rings(20000)
    g = World(sw.cfg)
    sw.add_to(g)
    bufs = [pinned_empty((sw.n, sw.cfg.num_variables, 4)) for _ in range(2)]
    expect = []
    for k in range(4):
        g.step()
        expect.append(g.read_beliefs(eta=False, lam=False, cov=False, valid=False)["mean"].copy())
    h = World(sw.cfg)
    sw.add_to(h)
    for k in range(4):
        h.step()
        h.read_means_into_async(bufs[k % 2])
        if k > 0:
            assert np.array_equal(bufs[(k - 1) % 2], expect[k - 1]), f"tick {k - 1}"
        h.reached_waypoint()  # uses the same device scratch: must not disturb the copy in flight
    h.readback_wait()
    assert np.array_equal(bufs[3 % 2], expect[3])
