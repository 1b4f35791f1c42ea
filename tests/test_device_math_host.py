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
            os.path.join(HERE, "..", "magics_b200", "csrc", "gbp_math.cuh")]
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
