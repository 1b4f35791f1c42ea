"""The oracle against every known-answer test the reference ships for the hot
path — CPU only."""
import json
import os

import numpy as np
import pytest

from oracle import oracle


def _load(golden_dir, name):
    return json.load(open(os.path.join(golden_dir, name)))


def test_schedule_sequences_match_reference_tests(golden_dir):
    cases = _load(golden_dir, "schedules.json")
    assert len(cases) >= 38
    for c in cases:
        oi, oe = oracle.schedule(c["kind"], c["internal"], c["external"])
        got = [[bool(a), bool(b)] for a, b in zip(oi, oe)]
        want = c["sequence"]
        assert got[: len(want)] == want, (c["schedule"], c["test"], c["internal"], c["external"], got)
        if c["exhaustive"]:
            assert len(got) == len(want), (c["schedule"], c["test"])
        assert len(got) == max(c["internal"], c["external"])


def test_variable_timesteps_match_reference_tests(golden_dir):
    for c in _load(golden_dir, "variable_timesteps.json"):
        got = oracle.variable_timesteps(c["lookahead_horizon"], c["lookahead_multiple"])
        assert got.tolist() == c["expected"]


def test_marginalise_block_extraction(golden_dir):
    g = _load(golden_dir, "marginalise.json")
    lam = np.array(g["matrix_8x8"])
    assert lam[0, 0] == 1 and lam[0, 4] == 17 and lam[4, 0] == 33 and lam[7, 7] == 64
    for idx, key in ((0, "blocks_marg_idx_0"), (4, "blocks_marg_idx_4")):
        aa, ab, ba, bb = oracle.extract_blocks(lam, idx)
        for got, name in ((aa, "aa"), (ab, "ab"), (ba, "ba"), (bb, "bb")):
            assert np.array_equal(got, np.array(g[key][name])), (idx, name)


def test_marginalise_single_neighbour_is_pass_through(golden_dir):
    g = _load(golden_dir, "marginalise.json")["single_neighbour"]
    eta, lam = np.array(g["eta"]), np.array(g["lam"])
    out = oracle.marginalise(eta, lam, 0)
    assert out is not None
    assert np.array_equal(out[0], eta) and np.array_equal(out[1], lam) and np.array_equal(out[2], np.zeros(4))


def test_marginalise_singular_block_is_empty(golden_dir):
    lam = np.array(_load(golden_dir, "marginalise.json")["matrix_8x8"])
    # the lower-right quadrant 49..64 has rank 2: `.inv()` is None -> Message::empty()
    assert oracle.marginalise(np.arange(8.0), lam, 0) is None


def test_inverse_contract():
    rng = np.random.default_rng(0)
    for _ in range(50):
        a = rng.normal(size=(4, 4))
        m = a @ a.T + 0.5 * np.eye(4)
        got = oracle.inv4(m)
        assert np.allclose(got, np.linalg.inv(m), rtol=1e-10, atol=1e-12)
    assert oracle.inv4(np.zeros((4, 4))) is None
    assert oracle.inv4(np.diag([1.0, 2.0, 0.0, 3.0])) is None
    d = oracle.inv4(np.eye(4) * 1e30)
    assert np.allclose(np.diag(d), 1e-30, rtol=1e-12)


def test_marginalise_schur_complement_against_numpy():
    rng = np.random.default_rng(1)
    for _ in range(20):
        a = rng.normal(size=(8, 8))
        lam = a @ a.T + np.eye(8)
        eta = rng.normal(size=8)
        for idx in (0, 4):
            a_sl = slice(idx, idx + 4)
            b_sl = slice(4, 8) if idx == 0 else slice(0, 4)
            bb_inv = np.linalg.inv(lam[b_sl, b_sl])
            want_eta = eta[a_sl] - lam[a_sl, b_sl] @ bb_inv @ eta[b_sl]
            want_lam = lam[a_sl, a_sl] - lam[a_sl, b_sl] @ bb_inv @ lam[b_sl, a_sl]
            got = oracle.marginalise(eta, lam, idx)
            assert np.allclose(got[0], want_eta, rtol=1e-10, atol=1e-10)
            assert np.allclose(got[1], want_lam, rtol=1e-10, atol=1e-10)
            assert np.array_equal(got[2], np.zeros(4))
