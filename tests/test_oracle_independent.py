"""Independent checks of the oracle's whole-iteration behaviour (CPU only).

The reference ships no test for FactorNode::update / the belief update / a GBP
iterate, so the oracle is additionally pinned against the
*mathematics*: on a tree (one robot, Dynamic factors only) synchronous GBP
converges to the exact marginals of the joint Gaussian, which numpy computes by
a dense solve written independently of the oracle's code path.
"""
import numpy as np

from magics_b200 import GbpConfig, scenarios
from oracle import oracle
from oracle.oracle import OracleWorld


def _dyn_potential(dt, sigma):
    I2, Z2 = np.eye(2), np.zeros((2, 2))
    qc_inv = np.eye(2) / sigma**2
    qi_inv = np.block([[12.0 / dt**3 * qc_inv, -6.0 / dt**2 * qc_inv], [-6.0 / dt**2 * qc_inv, 4.0 / dt * qc_inv]])
    J = np.block([[I2, dt * I2, -I2, Z2], [Z2, I2, Z2, -I2]])
    return J.T @ qi_inv @ J


def test_dynamics_chain_converges_to_joint_gaussian_marginals():
    sw = scenarios.circle(1)
    cfg = sw.cfg
    cfg.enable_obstacle = 0
    cfg.enable_interrobot = 0
    w = OracleWorld(cfg)
    sw.add_to(w)
    V = cfg.num_variables
    for _ in range(6):
        w.iterate()  # 60 synchronous sweeps >> tree diameter
    got = w.read_beliefs()

    t0 = np.float32(sw.radii[0]) / np.float32(2.0) / np.float32(cfg.target_speed)
    ts = sw.timesteps
    n = 4 * V
    lam = np.zeros((n, n))
    for i in range(V - 1):
        dt = float(np.float32(t0) * np.float32(ts[i + 1] - ts[i]))
        lam[4 * i:4 * i + 8, 4 * i:4 * i + 8] += _dyn_potential(dt, float(np.float32(cfg.sigma_factor_dynamics)))
    mu0, muN = sw.init_means[0, 0], sw.init_means[0, V - 1]
    # 1e30 priors pin the first and last variable: eliminate them
    free = np.arange(4, n - 4)
    fixed = np.r_[0:4, n - 4:n]
    xf = np.r_[mu0, muN]
    rhs = -lam[np.ix_(free, fixed)] @ xf
    x = np.linalg.solve(lam[np.ix_(free, free)], rhs)
    want = np.r_[mu0, x, muN].reshape(V, 4)
    assert np.allclose(got["mean"][0], want, rtol=1e-7, atol=1e-7)
    # marginal precision of an interior variable = Schur complement of the joint
    cov_joint = np.linalg.inv(lam[np.ix_(free, free)])
    for i in (1, V // 2, V - 2):
        k = 4 * (i - 1)
        assert np.allclose(got["cov"][0, i], cov_joint[k:k + 4, k:k + 4], rtol=1e-6, atol=1e-12)
    assert got["valid"].all()


def test_first_sweeps_follow_empty_message_rules():
    """Before information from the pinned ends arrives, interior beliefs stay at
    their initial mean (precision below the 1e-6 gate, variable.rs:276)."""
    sw = scenarios.circle(1)
    cfg = sw.cfg
    cfg.enable_obstacle = 0
    w = OracleWorld(cfg)
    sw.add_to(w)
    w.iterate_schedule([1], [0])
    b = w.read_beliefs()
    V = cfg.num_variables
    mid = V // 2
    assert np.array_equal(b["mean"][0, mid], sw.init_means[0, mid])
    assert np.abs(b["lam"][0, mid]).max() < 1e-6
    # the ends are pinned by the 1e30 prior
    assert np.allclose(np.diag(b["lam"][0, 0]), 1e30)


def test_two_robots_head_on_are_pushed_apart_symmetrically():
    cfg = GbpConfig(world_width=100.0, world_height=100.0)
    sw = scenarios.circle(2, circle_radius=6.0, cfg=cfg)
    w = OracleWorld(sw.cfg)
    sw.add_to(w)
    for _ in range(5):
        w.step()
    b = w.read_beliefs()
    off, nb, rn = w.read_connections()
    assert nb.tolist() == [1, 0] and rn.tolist() == [1, sw.cfg.num_variables]
    assert np.isfinite(b["mean"]).all()
    # InterRobot factors act: the planned paths leave the straight line y = 0
    assert np.abs(b["mean"][:, :, 1]).max() > 1e-3


def test_strict_reference_quirks_keeps_exactly_the_uncovered_pair():
    """SURVEY appendix B.1 in the oracle: four robots, {0, 3} and {1, 2} part; every robot loses two neighbours in one
    tick.  HashMap<RobotId, RobotId> keeps (0,2), (1,3), (2,3), (3,2): the pairs (0,2), (1,3), (2,3) are deleted in both
    directions and nobody deletes (0,1).  Without the quirk all four cross pairs go."""
    from dataclasses import replace

    from magics_b200 import scenarios
    from magics_b200.config import GbpConfig

    f32 = np.float32
    for strict in (0, 1):
        cfg = GbpConfig(target_speed=4.0, strict_reference_quirks=strict)
        ts = oracle.variable_timesteps(scenarios.lookahead_horizon(cfg.target_speed, 5.0), 3)
        cfg = replace(cfg, num_variables=int(ts.shape[0]))
        starts = np.array([[0, 0.75], [12, 0.75], [12, -0.75], [0, -0.75]], f32)
        far = np.array([[-14, 0.75], [26, 0.75], [26, -0.75], [-14, -0.75]], f32)
        sw = scenarios._finish(cfg, np.full(4, 0.3, f32), starts, far, ts, 5.0, sdf=scenarios.white_sdf())
        o = OracleWorld(sw.cfg)
        sw.add_to(o)
        for _ in range(12):
            o.step()
        off, nb, _ = o.read_connections()
        assert [sorted(nb[off[r]:off[r + 1]].tolist()) for r in range(4)] == [[3], [2], [1], [0]]
        sets = int(o.node_counts()[4]) // (cfg.num_variables - 1)
        assert sets == (6 if strict else 4)
        # robot 0 still holds the mirror message slot of robot 1's factor, and not of robot 2's
        has = lambda r, frm: o.has_mirror_slot(r, 1, frm)  # noqa: E731
        assert has(0, 1) == bool(strict) and has(1, 0) == bool(strict)
        assert not has(0, 2) and not has(2, 0) and not has(1, 3) and not has(3, 2)
