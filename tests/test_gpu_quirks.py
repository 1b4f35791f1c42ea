"""strict_reference_quirks: `delete_interrobot_factors` as the reference wrote it (robot.rs:1386-1439, SURVEY appendix
B.1).  Two pairs of robots part and meet again; every robot loses two neighbours in the same tick, so the lossy
per-robot map deletes the pairs (0,2), (3,1), (3,2) and (1,3), (2,0)... but nobody's LAST lost neighbour is the other end
of (0,1): that pair stays behind as a factor set nobody lists as a connection, keeps being iterated, and gets a second
set next to it when the robots meet again.  Engine == oracle (both in strict mode) at every tick."""
from dataclasses import replace

import numpy as np
import pytest

from magics_b200 import World, scenarios
from magics_b200.config import GbpConfig
from magics_b200.sharded import LocalShards
from oracle.oracle import OracleWorld
from tests.test_gpu_parity import check

pytestmark = pytest.mark.gpu
f32 = np.float32


def parting_pairs(strict: int):
    cfg = GbpConfig(target_speed=4.0, strict_reference_quirks=strict)
    ts = scenarios.get_variable_timesteps(scenarios.lookahead_horizon(cfg.target_speed, 5.0), 3)
    cfg = replace(cfg, num_variables=int(ts.shape[0]))
    y0 = 0.75
    starts = np.array([[0, y0], [12, y0], [12, -y0], [0, -y0]], f32)
    far = np.array([[-14, y0], [26, y0], [26, -y0], [-14, -y0]], f32)
    wps = [np.array([starts[k], far[k], starts[k]], f32) for k in range(4)]
    return scenarios._finish(cfg, np.full(4, 0.3, f32), starts, far, ts, 5.0, sdf=scenarios.white_sdf(),
                             name="parting-pairs", waypoints=wps)


def _sets(w, V):
    """(directed connections listed, directed InterRobot factor sets alive)"""
    return int(w.read_connections()[0][-1]), int(w.node_counts()[4]) // (V - 1)


def test_lossy_deletion_leaves_a_factor_set_behind_and_duplicates_it():
    sw = parting_pairs(1)
    V = sw.cfg.num_variables
    g, o = World(sw.cfg), OracleWorld(sw.cfg)
    sw.add_to(g)
    sw.add_to(o)
    seen = set()
    for tick in range(70):
        rg = g.reached_waypoint((1, 0, 1, 1.0), (0, 0, 1, 1.0))
        ro = o.reached_waypoint((1, 0, 1, 1.0), (0, 0, 1, 1.0))
        assert np.array_equal(rg, ro)
        g.step()
        o.step()
        check(g, o, f"strict tick {tick}")
        assert _sets(g, V) == _sets(o, V)
        seen.add(_sets(g, V))
    # together (12 = 4 robots x 3), parted with the pair (0,1) left behind in both directions, together again
    # with a second set between 0 and 1
    assert {(12, 12), (4, 6), (12, 14)} <= seen, seen
    off, nb, rn = g.read_connections()
    assert np.array_equal(np.diff(off), [3, 3, 3, 3]) and sorted(nb[off[0]:off[1]].tolist()) == [1, 2, 3]


def test_default_mode_deletes_every_lost_pair():
    sw = parting_pairs(0)
    V = sw.cfg.num_variables
    g, o = World(sw.cfg), OracleWorld(sw.cfg)
    sw.add_to(g)
    sw.add_to(o)
    seen = set()
    for tick in range(60):
        for w in (g, o):
            w.reached_waypoint((1, 0, 1, 1.0), (0, 0, 1, 1.0))
            w.step()
        if tick % 6 == 0:
            check(g, o, f"default tick {tick}")
        seen.add(_sets(g, V))
    assert seen == {(12, 12), (4, 4)}, seen


def test_strict_mode_is_refused_on_sharded_worlds():
    sw = parting_pairs(1)
    with pytest.raises(RuntimeError, match="single-GPU"):
        LocalShards(sw.cfg, 2)
