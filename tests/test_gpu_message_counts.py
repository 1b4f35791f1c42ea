"""MessageCount (factorgraph/mod.rs:103-137; FactorGraph::messages_sent / messages_received, read by export.rs:434-439):
the engine's accounting kernels against the oracle, which increments a counter at every site the reference does
(variable.rs:186-188, 332; factor/mod.rs:313-315, 367, 452).  Exact integers, compared after every tick."""
import numpy as np
import pytest

from magics_b200 import World, scenarios
from magics_b200.config import FACTOR_DYNAMIC, FACTOR_INTERROBOT, FACTOR_OBSTACLE
from magics_b200.sharded import LocalShards
from oracle.oracle import OracleWorld
from tests.test_gpu_parity import check
from tests.test_gpu_quirks import parting_pairs

pytestmark = pytest.mark.gpu


def _pair(sw):
    g, o = World(sw.cfg), OracleWorld(sw.cfg)
    g.set_message_counting(True)
    sw.add_to(g)
    sw.add_to(o)
    return g, o


def _same_counts(g, o, what):
    cg, co = g.read_message_counts(), o.read_message_counts()
    assert np.array_equal(cg, co), f"{what}: first difference at robot {np.argwhere(cg != co)[0][0]}: {cg[cg != co][:4]} vs {co[cg != co][:4]}"


def test_counts_follow_every_half_step_of_one_tick():
    sw = scenarios.circle(6, circle_radius=8.0)
    g, o = _pair(sw)
    _same_counts(g, o, "creation")
    for name in ("update_topology", "update_prior_of_horizon_state", "update_prior_of_current_state"):
        for w in (g, o):
            getattr(w, name)()
        _same_counts(g, o, name)
    for half in ("int", "ext", "int", "int", "ext", "ext"):
        for w in (g, o):
            if half == "int":
                w.internal_factor_iteration()
                w.internal_variable_iteration()
            else:
                w.external_factor_iteration()
                w.external_variable_iteration()
        _same_counts(g, o, half)
    c = g.read_message_counts()
    assert (c > 0).all()


def test_crossing_circle_with_radio_failures_idle_robots_and_despawn():
    sw = scenarios.circle(14, circle_radius=16.0)
    g, o = _pair(sw)
    rng = np.random.default_rng(11)
    for tick in range(60):
        ant = (rng.random(sw.n) > (0.25 if tick >= 4 else 0.0)).astype(np.uint8)
        idle = (rng.random(sw.n) < (0.1 if 10 <= tick < 20 else 0.0)).astype(np.uint8)
        for w in (g, o):
            w.set_comms(ant, idle)
            if tick == 30:
                w.remove_robots([3, 9])
            w.step()
        _same_counts(g, o, f"tick {tick}")
        if tick % 15 == 0:
            check(g, o, f"tick {tick}")
    assert (g.read_message_counts()[[3, 9]] == 0).all()  # the graphs went with the robots


def test_tracking_factors_count_from_the_tenth_factor_iteration_and_toggles():
    sw = scenarios.junction_twoway(per_lane=1)
    g, o = _pair(sw)
    V = sw.cfg.num_variables
    for tick in range(8):
        for w in (g, o):
            if tick == 3:
                w.change_factor_enabled(FACTOR_OBSTACLE, 0)
            if tick == 4:
                w.change_factor_enabled(FACTOR_INTERROBOT, 0)
            if tick == 5:
                w.change_factor_enabled(FACTOR_INTERROBOT, 1)
                w.change_factor_enabled(FACTOR_OBSTACLE, 1)
            if tick == 6:
                w.change_factor_enabled(FACTOR_DYNAMIC, 0)
            if tick == 7:
                w.change_factor_enabled(FACTOR_DYNAMIC, 1)
                means = o.read_beliefs()["mean"].reshape(sw.n, V, 4)
                w.change_prior_of_variable(4, [0, 2, 5], means[[0, 2, 5], 4] + 0.25)
                w.change_prior_of_variable(0, [1], means[[1], 0])
            w.step()
        _same_counts(g, o, f"tick {tick}")
        if tick == 4:
            # beliefs are compared up to the first RE-enabling: a factor that is switched back on resumes in the
            # reference with the inbox it had when it was switched off (receive_message_from ignores a disabled
            # factor, factor/mod.rs:308-310), in the engine from the current records
            check(g, o, "before re-enabling")


def test_counts_with_zombie_and_duplicate_factor_sets():
    sw = parting_pairs(1)
    g, o = _pair(sw)
    for tick in range(55):
        for w in (g, o):
            w.reached_waypoint((1, 0, 1, 1.0), (0, 0, 1, 1.0))
            w.step()
        _same_counts(g, o, f"strict tick {tick}")


def test_counting_is_off_by_default_and_refused_on_shards():
    sw = scenarios.circle(4, circle_radius=6.0)
    g = World(sw.cfg)
    sw.add_to(g)
    with pytest.raises(RuntimeError, match="counting is off"):
        g.read_message_counts()
    with pytest.raises(RuntimeError, match="before the first robot"):
        g.set_message_counting(True)
    sh = LocalShards(sw.cfg, 2)
    with pytest.raises(RuntimeError, match="single-GPU"):
        sh.shards[0].set_message_counting(True)
    sh.close()
