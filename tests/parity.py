"""Parity metric shared by the tests, smoke() and bench.py.

North star: beliefs and means match the reference's CPU GBP within 1e-9
relative in f64.  "Relative" is taken per 4-vector / 4x4 block against the
block's own magnitude with a floor of 1 (metres, m/s for means; for precision
and information blocks the floor is far below any real entry, >= 1e2 for every
shipped sigma, and above the reference's own "precision is zero" threshold of
1e-6, variable.rs:276, under which a block is round-off noise by the
reference's own definition).
"""
import numpy as np

RTOL = 1e-9


def block_rel_err(got: np.ndarray, ref: np.ndarray, block_axes: int) -> float:
    """max over blocks of max|got-ref| / max(1, max|ref|); non-finite entries must match exactly."""
    got = np.asarray(got, np.float64)
    ref = np.asarray(ref, np.float64)
    assert got.shape == ref.shape, (got.shape, ref.shape)
    fin = np.isfinite(ref)
    if not np.array_equal(fin, np.isfinite(got)):
        return float("inf")
    if not np.array_equal(np.where(fin, 0.0, ref), np.where(fin, 0.0, got), equal_nan=True):
        return float("inf")
    g = np.where(fin, got, 0.0)
    r = np.where(fin, ref, 0.0)
    axes = tuple(range(ref.ndim - block_axes, ref.ndim))
    scale = np.maximum(1.0, np.abs(r).max(axis=axes, keepdims=True))
    return float((np.abs(g - r) / scale).max()) if ref.size else 0.0


def cov_rel_err(got, ref) -> float:
    """Covariances span 1e-30 (fixed poses) .. 1e2; compare each 4x4 block against its own max."""
    got = np.asarray(got, np.float64)
    ref = np.asarray(ref, np.float64)
    fin = np.isfinite(ref) & np.isfinite(got)
    if not np.array_equal(np.isfinite(ref), np.isfinite(got)):
        return float("inf")
    g = np.where(fin, got, 0.0)
    r = np.where(fin, ref, 0.0)
    scale = np.abs(r).max(axis=(-1, -2), keepdims=True)
    scale = np.where(scale > 0, scale, 1.0)
    return float((np.abs(g - r) / scale).max()) if ref.size else 0.0


def compare_beliefs(got: dict, ref: dict) -> dict:
    out = {
        "mean": block_rel_err(got["mean"], ref["mean"], 1),
        "eta": block_rel_err(got["eta"], ref["eta"], 1),
        "lam": block_rel_err(got["lam"], ref["lam"], 2),
        "cov": cov_rel_err(got["cov"], ref["cov"]),
        "valid": 0.0 if np.array_equal(got["valid"], ref["valid"]) else float("inf"),
    }
    return out


def assert_beliefs_match(got: dict, ref: dict, rtol: float = RTOL, what: str = ""):
    errs = compare_beliefs(got, ref)
    bad = {k: v for k, v in errs.items() if not v <= rtol}
    assert not bad, f"{what} parity broken (rtol {rtol}): {errs}"
    return errs
