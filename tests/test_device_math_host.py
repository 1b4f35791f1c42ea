"""The DEVICE math header (magics_b200/csrc/gbp_math.cuh) compiled for the host by g++ through
tests/host_math/host_math_shim.h, checked against the oracle and IEEE division without a GPU.

* inv4 general expansion: bit-identical to the oracle's restatement of ndarray-inverse `inv()`.
* inv4 decoupled path (x and y chains do not mix): every value equal to the oracle's; only the sign
  of exact zeros in the structurally-zero entries may differ (documented in gbp_math.cuh).
* divide_all (shared-divisor division): bit-identical to N plain divisions, extreme operands included.
"""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

from oracle import oracle as oo

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.path.join(HERE, "host_math", "host_math.cpp")
P = C.POINTER(C.c_double)


def _build(tag: str, flags: list[str]) -> C.CDLL:
    out_dir = os.path.join(HERE, "host_math", "_build")
    os.makedirs(out_dir, exist_ok=True)
    out = os.path.join(out_dir, f"libhost_math_{tag}.so")
    deps = [SRC, os.path.join(HERE, "host_math", "host_math_shim.h"),
            os.path.join(HERE, "..", "magics_b200", "csrc", "gbp_math.cuh"),
            os.path.join(HERE, "..", "magics_b200", "csrc", "gbp_math_axis.cuh"),
            os.path.join(HERE, "..", "magics_b200", "csrc", "gbp_collide.cuh")]
    if not os.path.exists(out) or any(os.path.getmtime(out) < os.path.getmtime(d) for d in deps):
        subprocess.run(["g++", "-O2", "-std=c++17", "-ffp-contract=off", "-fPIC", "-shared",
                        "-I", os.path.join(HERE, "host_math")] + flags + ["-o", out, SRC], check=True)
    return C.CDLL(out)


@pytest.fixture(scope="module")
def libs():
    return _build("default", []), _build("general", ["-DGBP_INV4_DECOUPLED=0"]), oo.lib()


def _inv(lib, fn, m):
    m = np.ascontiguousarray(m, dtype=np.float64).reshape(16)
    o = np.zeros(16)
    r = getattr(lib, fn)(m.ctypes.data_as(P), o.ctypes.data_as(P))
    return r, o


def _matrices(n, seed):
    rng = np.random.default_rng(seed)
    for t in range(n):
        kind = t % 5
        vals = rng.standard_normal((4, 4)) * 10.0 ** rng.integers(-8, 31, size=(4, 4))
        if kind == 4:
            yield vals  # x and y coupled: general path
            continue
        m = np.where((np.add.outer(np.arange(4), np.arange(4)) % 2) == 0, vals, 0.0)
        if kind == 1:  # symmetric, diagonally dominant (a precision matrix)
            m = (m + m.T) / 2 + 2 * np.diag(np.abs(np.diag(m)))
        elif kind == 2:  # zero rows / a 1e30 prior
            m[rng.integers(0, 4), :] *= float(rng.random() > 0.5)
            m[0, 0] = 1e30
        elif kind == 3:  # negative zeros in the structurally-zero places, singular blocks
            m = np.where(m == 0.0, -0.0, m)
            if rng.random() < 0.3:
                m[2, :] = m[0, :]
        yield m


def test_inv4_general_path_is_bit_identical_to_the_oracle(libs):
    _, gen, orc = libs
    for m in _matrices(20000, 1):
        r0, o0 = _inv(orc, "gbpo_inv4", m)
        r1, o1 = _inv(gen, "hm_inv4", m)
        assert r0 == r1
        if r0:
            assert np.array_equal(o0.view(np.uint64), o1.view(np.uint64)), (m, o0, o1)


def test_inv4_decoupled_path_matches_every_value(libs):
    dev, _, orc = libs
    sign_only = 0
    for m in _matrices(40000, 2):
        r0, o0 = _inv(orc, "gbpo_inv4", m)
        r1, o1 = _inv(dev, "hm_inv4", m)
        assert r0 == r1, m
        if not r0:
            continue
        assert np.array_equal(o0, o1, equal_nan=True), (m, o0, o1)
        diff = o0.view(np.uint64) != o1.view(np.uint64)
        if diff.any():
            sign_only += 1
            assert np.all(o0[diff] == 0.0)  # only exact zeros may differ (in sign)
    assert sign_only > 0  # the fast path was exercised


def test_inv4_non_finite_entries_take_the_general_path(libs):
    dev, _, orc = libs
    m = np.diag([np.inf, 2.0, 3.0, 4.0])
    r0, o0 = _inv(orc, "gbpo_inv4", m)
    r1, o1 = _inv(dev, "hm_inv4", m)
    assert r0 == r1 and np.array_equal(o0, o1, equal_nan=True)
    m = np.diag([1.0, 2.0, np.nan, 4.0])
    r0, o0 = _inv(orc, "gbpo_inv4", m)
    r1, o1 = _inv(dev, "hm_inv4", m)
    assert r0 == r1 and np.array_equal(o0, o1, equal_nan=True)


def test_divide_all_is_bit_identical_to_plain_division(libs):
    dev, _, _ = libs
    rng = np.random.default_rng(3)
    specials = np.array([0.0, -0.0, 1e-310, -1e-310, 1e300, -1e300, np.inf, -np.inf, np.nan, 1.0, 3.0, 1e-200])
    for t in range(5000):
        c = rng.standard_normal(16) * 10.0 ** rng.integers(-160, 160, size=16)
        det = float(rng.standard_normal() * 10.0 ** rng.integers(-160, 160))
        if t % 3 == 0:
            c[rng.integers(0, 16, size=4)] = rng.choice(specials, size=4)
        if t % 7 == 0:
            det = float(rng.choice(specials[2:]))
        o = np.zeros(16)
        dev.hm_divide_all16(c.ctypes.data_as(P), C.c_double(det), o.ctypes.data_as(P))
        with np.errstate(all="ignore"):
            ref = c / det
        same = (o.view(np.uint64) == ref.view(np.uint64)) | (np.isnan(o) & np.isnan(ref))
        assert same.all(), (c, det, o, ref)


# ---- the device factor / belief arithmetic against independent float64 linear algebra -------------------
def _spd(rng, scale):
    a = rng.standard_normal((4, 4))
    return (a @ a.T + 4 * np.eye(4)) * scale


def _schur(eta, lam, keep):
    """Marginalise the other 4-block out of an 8-dim information form (marginalise_factor_distance.rs:55-127)."""
    a = slice(0, 4) if keep == 0 else slice(4, 8)
    b = slice(4, 8) if keep == 0 else slice(0, 4)
    lbb_inv = np.linalg.inv(lam[b, b])
    return eta[a] - lam[a, b] @ lbb_inv @ eta[b], lam[a, a] - lam[a, b] @ lbb_inv @ lam[b, a]


def _rel(got, ref):
    return float(np.max(np.abs(got - ref)) / max(1e-300, np.max(np.abs(ref))))


def test_dynamic_factor_message_matches_the_schur_complement(libs):
    dev, _, _ = libs
    rng = np.random.default_rng(7)
    I2, Z2 = np.eye(2), np.zeros((2, 2))
    for trial in range(300):
        dt, sigma = float(rng.uniform(0.05, 2.0)), float(rng.choice([0.1, 1.0]))
        qs = 1.0 / sigma ** 2
        # DynamicFactor (factor/dynamic.rs:22-52): h = J x, z = 0, so eta_p = 0 and Lambda_p = J^T Qi^-1 J
        J = np.block([[I2, dt * I2, -I2, Z2], [Z2, I2, Z2, -I2]])
        Qi_inv = qs * np.block([[12 / dt ** 3 * I2, -6 / dt ** 2 * I2], [-6 / dt ** 2 * I2, 4 / dt * I2]])
        lam_p = J.T @ Qi_inv @ J
        for keep in (0, 1):
            for nonempty in (0, 1):
                oe, ol = rng.standard_normal(4) * 10, _spd(rng, float(10.0 ** rng.integers(-1, 4)))
                other = np.concatenate([oe, ol.reshape(-1)])
                eta, lam = np.zeros(4), np.zeros(16)
                ok = dev.hm_dyn_message(keep, C.c_double(dt), C.c_double(qs), nonempty, other.ctypes.data_as(P),
                                        eta.ctypes.data_as(P), lam.ctypes.data_as(P))
                assert ok == 1
                e8, l8 = np.zeros(8), lam_p.copy()
                o = slice(4, 8) if keep == 0 else slice(0, 4)  # the OTHER variable's block
                if nonempty:
                    e8[o] += oe
                    l8[o, o] += ol
                re, rl = _schur(e8, l8, keep)
                # against the size of the unreduced block: the potential alone marginalises to (almost) nothing,
                # a difference of two equal matrices, so the result's own size is round-off
                k = slice(0, 4) if keep == 0 else slice(4, 8)
                scale = np.abs(l8[k, k]).max()
                assert np.max(np.abs(lam.reshape(4, 4) - rl)) <= 1e-9 * scale
                assert np.max(np.abs(eta - re)) <= 1e-9 * max(1.0, np.abs(re).max(), np.abs(e8).max())


def test_belief_moments_match_a_dense_solve(libs):
    dev, _, _ = libs
    rng = np.random.default_rng(8)
    for trial in range(500):
        lam = _spd(rng, float(10.0 ** rng.integers(-2, 6)))
        if trial % 3 == 0:  # x and y decoupled, a 1e30 prior
            lam[np.add.outer(np.arange(4), np.arange(4)) % 2 == 1] = 0.0
            lam[0, 0] += 1e30
        eta = rng.standard_normal(4) * np.abs(lam).max()
        mu, cov, valid = np.zeros(4), np.zeros(16), C.c_int(0)
        taken = dev.hm_belief_moments(eta.ctypes.data_as(P), lam.reshape(-1).copy().ctypes.data_as(P),
                                      mu.ctypes.data_as(P), cov.ctypes.data_as(P), C.byref(valid))
        assert taken == 1 and valid.value == 1
        ref = np.linalg.inv(lam)
        assert _rel(cov.reshape(4, 4), ref) < 1e-9
        assert np.max(np.abs(mu - ref @ eta)) <= 1e-9 * max(1.0, np.abs(ref @ eta).max())
    # all-zero precision (an interior variable before any message): nothing is taken, mean untouched
    mu = np.array([1.0, 2.0, 3.0, 4.0])
    z = np.zeros(16)
    assert dev.hm_belief_moments(np.zeros(4).ctypes.data_as(P), z.ctypes.data_as(P), mu.ctypes.data_as(P),
                                 np.zeros(16).ctypes.data_as(P), C.byref(C.c_int(0))) == 0
    assert np.array_equal(mu, [1.0, 2.0, 3.0, 4.0])


def test_interrobot_message_matches_the_linearised_factor(libs):
    dev, _, _ = libs
    rng = np.random.default_rng(9)
    n_msg = n_skip = 0
    for trial in range(2000):
        a_first = int(rng.integers(0, 2))
        dsafe, sigma = float(rng.uniform(1.0, 4.0)), 0.01
        lm = 1.0 / sigma ** 2
        mu_a = rng.uniform(-3, 3, 2)
        mu_b = mu_a + rng.uniform(-1, 1, 2) * dsafe * 1.2
        tiny = 9.999999974752427e-7 * float(rng.integers(1, 50))
        a_nonempty = int(rng.integers(0, 2))
        ea, la = rng.standard_normal(4) * 5, _spd(rng, float(10.0 ** rng.integers(0, 4)))
        rec = np.concatenate([ea, la.reshape(-1)])
        eta2, lam4 = np.zeros(2), np.zeros(4)
        ok = dev.hm_interrobot_message(a_first, mu_a.ctypes.data_as(P), mu_b.ctypes.data_as(P), a_nonempty,
                                       rec.ctypes.data_as(P), C.c_double(dsafe), C.c_double(tiny), C.c_double(lm),
                                       eta2.ctypes.data_as(P), lam4.ctypes.data_as(P))
        skip = dev.hm_interrobot_skip(a_first, mu_a.ctypes.data_as(P), mu_b.ctypes.data_as(P), C.c_double(dsafe))
        x0, x1 = (mu_a, mu_b) if a_first else (mu_b, mu_a)  # slot order by robot id (id.rs:83-118)
        # InterRobotFactor::skip (interrobot.rs:213-226): squared distance without the tiny offset
        assert skip == int(float(np.sum((x0 - x1) ** 2)) >= dsafe * dsafe)
        if skip:
            assert ok == 0
            n_skip += 1
            continue
        # measure / jacobian (interrobot.rs:121-204) at x = (x0, 0, 0, x1, 0, 0)
        d = x0 - x1 + tiny
        r = float(np.linalg.norm(d))
        J = np.zeros(8)
        h = 0.0
        if r <= dsafe:
            h = 1.0 - r / dsafe
            J[0:2] = -d / (dsafe * r)
            J[4:6] = d / (dsafe * r)
        x = np.concatenate([x0, [0, 0], x1, [0, 0]])
        v0 = J @ x + (0.0 - h)
        eta8, lam8 = J * lm * v0, np.outer(J, J) * lm
        sa = slice(0, 4) if a_first else slice(4, 8)  # A's (the factor owner's) variable
        if a_nonempty:
            eta8[sa] += ea
            lam8[sa, sa] += la
        if not a_nonempty:
            # the potential alone is rank one: Lambda of A's block is singular, the reference returns Empty
            assert ok == 0
            continue
        re, rl = _schur(eta8, lam8, keep=1 if a_first else 0)
        assert ok == 1
        n_msg += 1
        assert np.max(np.abs(eta2 - re[:2])) <= 1e-9 * max(1.0, np.abs(re).max())
        assert np.max(np.abs(lam4.reshape(2, 2) - rl[:2, :2])) <= 1e-9 * max(1.0, np.abs(rl).max())
        assert np.max(np.abs(rl[2:, :])) <= 1e-9 * max(1.0, np.abs(rl).max())  # the rest of the message is zero
    assert n_msg > 200 and n_skip > 200


# ---- gbp_math_axis.cuh: the two-lanes-per-variable arithmetic of the decoupled regime, bit for bit against
# ---- the general 4x4 routines of gbp_math.cuh (which the tests above tie to the oracle) -------------------------
AX = [np.array([0, 2, 8, 10]), np.array([5, 7, 13, 15])]   # row-major 4x4 indices of each axis' block
AXV = [np.array([0, 2]), np.array([1, 3])]                  # vector components of each axis


def _bits_equal(a, b):
    a, b = np.asarray(a, np.float64), np.asarray(b, np.float64)
    return bool(np.all((a.view(np.uint64) == b.view(np.uint64)) | (np.isnan(a) & np.isnan(b))))


def _decoupled(rng, kind):
    """A 4x4 with exact zeros wherever row + column is odd."""
    vals = rng.standard_normal((4, 4)) * 10.0 ** rng.integers(-6, 12, size=(4, 4))
    m = np.where((np.add.outer(np.arange(4), np.arange(4)) % 2) == 0, vals, 0.0)
    if kind == 1:
        m = (m + m.T) / 2 + 2 * np.diag(np.abs(np.diag(m)))
    elif kind == 2:
        m[0, 0] = 1e30
        m[1, 1] = 1e30
    elif kind == 3 and rng.random() < 0.5:
        m[2, :] = m[0, :]  # singular x block
    return m


def test_inv_axis_has_the_bits_of_inv4(libs):
    dev, _, _ = libs
    rng = np.random.default_rng(21)
    taken = 0
    for t in range(20000):
        m = _decoupled(rng, t % 4).reshape(16)
        if t % 97 == 0:
            m[rng.choice(np.concatenate(AX))] = rng.choice([np.inf, -np.inf, np.nan])
        r0, o0 = _inv(dev, "hm_inv4", m)
        det_fin = True
        outs = []
        for a in (0, 1):
            P_, Q_ = m[AX[a]].copy(), m[AX[1 - a]].copy()
            O = np.zeros(4)
            ok = dev.hm_inv_axis(a, P_.ctypes.data_as(P), Q_.ctypes.data_as(P), O.ctypes.data_as(P))
            outs.append((ok, O))
        assert outs[0][0] == outs[1][0]
        if outs[0][0]:
            # the lane pair took the update: inv4 must have taken it too, with the same bits in both blocks
            assert r0 == 1
            for a in (0, 1):
                assert _bits_equal(outs[a][1], o0[AX[a]]), (m, outs[a][1], o0[AX[a]])
            taken += 1
        else:
            # refused: singular, or a non-finite determinant (inv4's general path)
            assert r0 == 0 or not np.all(np.isfinite(m))
    assert taken > 10000


def test_belief_axis_has_the_bits_of_belief_moments(libs):
    dev, _, _ = libs
    rng = np.random.default_rng(22)
    taken = 0
    for t in range(5000):
        lam = _decoupled(rng, 1 if t % 3 else 2).reshape(16)
        if t % 11 == 0:
            lam *= 1e-9  # nothing above the 1e-6 threshold of variable.rs:276
        eta = rng.standard_normal(4) * 10.0 ** rng.integers(-3, 8, size=4)
        mu0 = rng.standard_normal(4)
        mu, cov, valid = mu0.copy(), np.zeros(16), C.c_int(0)
        tk = dev.hm_belief_moments(eta.ctypes.data_as(P), lam.copy().ctypes.data_as(P), mu.ctypes.data_as(P),
                                   cov.ctypes.data_as(P), C.byref(valid))
        oks = []
        for a in (0, 1):
            e2, m2 = eta[AXV[a]].copy(), mu0[AXV[a]].copy()
            P_, Q_ = lam[AX[a]].copy(), lam[AX[1 - a]].copy()
            ok = dev.hm_belief_axis(a, e2.ctypes.data_as(P), P_.ctypes.data_as(P), Q_.ctypes.data_as(P), m2.ctypes.data_as(P))
            oks.append(ok)
            if ok:
                assert tk == 1 and valid.value == 1
                assert _bits_equal(m2, mu[AXV[a]]), (lam, eta, m2, mu)
        if tk == 1 and valid.value == 1:
            assert oks == [1, 1]
            taken += 1
        else:
            assert oks == [0, 0]
    assert taken > 3000


def test_dyn_message_axis_has_the_bits_of_dyn_message(libs):
    dev, _, _ = libs
    rng = np.random.default_rng(23)
    n_ok = 0
    for t in range(6000):
        dt = float(np.float32(rng.uniform(0.02, 2.5)))
        qs = 1.0 / float(rng.choice([0.1, 1.0, 0.5])) ** 2
        keep, nonempty = int(rng.integers(0, 2)), int(t % 5 != 0)
        ol = _decoupled(rng, 1 if t % 7 else 0).reshape(16)
        oe = rng.standard_normal(4) * 10.0 ** rng.integers(-3, 6, size=4)
        other = np.concatenate([oe, ol])
        eta, lam = np.zeros(4), np.zeros(16)
        ok = dev.hm_dyn_message(keep, C.c_double(dt), C.c_double(qs), nonempty, other.ctypes.data_as(P),
                                eta.ctypes.data_as(P), lam.ctypes.data_as(P))
        oks = []
        for a in (0, 1):
            e2, P_, Q_ = oe[AXV[a]].copy(), ol[AX[a]].copy(), ol[AX[1 - a]].copy()
            re, rl = np.zeros(2), np.zeros(4)
            oka = dev.hm_dyn_message_axis(keep, a, C.c_double(dt), C.c_double(qs), nonempty, e2.ctypes.data_as(P),
                                          P_.ctypes.data_as(P), Q_.ctypes.data_as(P), re.ctypes.data_as(P),
                                          rl.ctypes.data_as(P))
            oks.append(oka)
            if oka:
                assert ok == 1
                assert _bits_equal(re, eta[AXV[a]]), (t, a, re, eta)
                assert _bits_equal(rl, lam[AX[a]]), (t, a, rl, lam)
        if ok and oks == [1, 1]:
            # the entries outside the two blocks are exact zeros in the general result
            rest = np.setdiff1d(np.arange(16), np.concatenate(AX))
            assert np.all(lam[rest] == 0.0)
            n_ok += 1
        else:
            assert oks[0] == oks[1] or ok == 0
    assert n_ok > 4000


# ---- gbp_collide.cuh: the environment-collision predicate of the engine against the oracle's restatement of
# parry2d's intersection_test — two independent sources, every shape kind, bit-equal booleans incl. rim points
def test_collider_predicate_of_the_engine_equals_the_oracle_restatement(libs):
    dev, _, _ = libs
    from magics_b200 import scenarios
    from magics_b200.environment import Collider
    from oracle.oracle import OracleWorld

    rng = np.random.default_rng(21)
    R = np.float32(0.6)
    hexagon = tuple((float(np.float32(1.5 * np.cos(k * np.pi / 3))), float(np.float32(1.5 * np.sin(k * np.pi / 3))))
                    for k in range(6))
    cols = [Collider("ball", (3.0, -2.0), 0.0, radius=1.25),
            Collider("cuboid", (-4.0, 1.0), 0.7, half_extents=(2.0, 0.5)),
            Collider("cuboid", (0.0, -6.0), 0.0, half_extents=(3.0, 1.0)),
            Collider("triangle", (0.5, 5.0), 0.7, points=((-1.0, -0.5), (2.0, -0.5), (0.3, 1.7))),
            Collider("triangle", (5.0, -5.0), -2.1, points=((2.0, -0.5), (-1.0, -0.5), (0.3, 1.7))),  # clockwise
            Collider("convex-polygon", (6.0, 6.0), -0.7, points=hexagon)]
    pts = rng.uniform(-9, 11, size=(20000, 2)).astype(np.float32)
    # rim: points at exactly the robot radius from a face / corner / ball, and their f32 neighbours
    rim = [(3.0 + 1.25 + R, -2.0), (3.0 + R, -6.0), (0.0, -5.0 + R), (-3.0 - R, -7.0 - R), (3.0, -5.0 + R)]
    k = 0
    for x, z in rim:
        for dx in (-1, 0, 1):
            pts[k] = (np.nextafter(np.float32(x), np.float32(np.inf * dx)) if dx else np.float32(x), np.float32(z))
            k += 1
    sw = scenarios.circle(len(pts), 10.0, robot_radius=float(R))
    sw.positions[:] = pts
    o = OracleWorld(sw.cfg)
    sw.add_to(o)
    kinds = {"ball": 0, "cuboid": 1, "triangle": 2, "convex-polygon": 3}
    F = C.POINTER(C.c_float)
    for c in cols:
        o.set_environment_colliders([c])
        o.update_environment_collisions()
        want = o.read_environment_collisions().astype(np.uint8)
        verts = np.ascontiguousarray(c.points if c.points else [(0.0, 0.0)], np.float32)
        got = np.zeros(len(pts), np.uint8)
        a = np.float32(c.angle)
        dev.hm_collider_hits(kinds[c.kind], C.c_float(c.translation[0]), C.c_float(c.translation[1]),
                             C.c_float(float(np.cos(a))), C.c_float(float(np.sin(a))), C.c_float(c.radius),
                             C.c_float(c.half_extents[0]), C.c_float(c.half_extents[1]), len(c.points),
                             verts.ctypes.data_as(F), len(pts), pts.ctypes.data_as(F), C.c_float(float(R)),
                             got.ctypes.data_as(C.POINTER(C.c_ubyte)))
        assert np.array_equal(got, want), (c.kind, int((got != want).sum()))
        assert 10 < want.sum() < len(pts) - 10
