"""How far can a reference binary on FMA hardware sit from the oracle's bits?  

The reference's matrix-matrix products (J^T Lambda J, Lambda_ab Lambda_bb^-1 Lambda_ba) go through the `matrixmultiply`
crate, whose x86-64 kernels may fuse multiply-adds at run time; the oracle (and the engine, -fmad=false) round every
product.  No reference binary can be built here, so this test bounds the effect from the other side: the SAME oracle
source compiled with its matrix-matrix products accumulated by std::fma (`make -C oracle fma`, -DGBPO_FMA_MATMUL: the
k-ascending fused accumulation of a dgemm micro-kernel; Rust itself never contracts, so nothing else in the reference
can fuse) is run beside the uncontracted build on scenarios that exercise every factor kind, and the two must agree
within the north star's 1e-9 relative on beliefs and means and exactly on all graph indexing.  The fused build is never
the parity checker.
"""
import numpy as np
import pytest

from magics_b200 import scenarios
from oracle.oracle import OracleWorld
from tests.parity import RTOL, compare_beliefs


def _has_fma() -> bool:
    try:
        with open("/proc/cpuinfo") as f:
            return any(" fma " in line for line in f if line.startswith("flags"))
    except OSError:
        return False


pytestmark = pytest.mark.skipif(not _has_fma(), reason="host CPU has no FMA unit")


def _run_pair(sw, ticks):
    """Worst block-relative belief difference (tests/parity.py) between the two builds over `ticks` ticks."""
    a, b = OracleWorld(sw.cfg), OracleWorld(sw.cfg, fma=True)
    sw.add_to(a)
    sw.add_to(b)
    worst = 0.0
    for _ in range(ticks):
        a.step()
        b.step()
        worst = max(worst, max(compare_beliefs(b.read_beliefs(), a.read_beliefs()).values()))
    same_graph = all(np.array_equal(x, y) for x, y in zip(a.read_connections(), b.read_connections()))
    return worst, same_graph


# scenario -> (generator, ticks, bound).  Measured on this image (gcc 13, x86-64 FMA3): 6.6e-13, 1.3e-11, 5.0e-12, 4.6e-9.
CASES = {
    # ten robots far apart on a 50 m circle: Dynamic factors and belief updates only for the first 20 ticks
    "apart": (lambda: scenarios.circle(10, circle_radius=50.0), 20, 1e-11),
    "single": (lambda: scenarios.circle(1, circle_radius=12.0), 20, 1e-10),
    # Obstacle factors on a sloping SDF, InterRobot factors between passing robots
    "complex": (lambda: scenarios.complex_environment(n=8, seed=2), 40, 1e-10),
    # three robots meeting in the middle: the crossing amplifies the last-bit differences ~1e4-fold
    "crossing3": (lambda: scenarios.circle(3, circle_radius=12.0), 20, 1e-7),
}


@pytest.mark.parametrize("name", sorted(CASES))
def test_fused_matrix_products_move_beliefs_by_rounding_only(name):
    make, ticks, bound = CASES[name]
    worst, same_graph = _run_pair(make(), ticks)
    print(f"fma sensitivity {name}: {worst:.3g}")
    assert same_graph, "graph indexing must not depend on how products are rounded"
    assert 0.0 < worst, "the fused build reproduced every bit - not fused?"
    assert worst <= bound, f"{name}: fused vs rounded matrix products differ by {worst:.3g}"
    if name != "crossing3":
        assert worst <= RTOL


def test_a_symmetric_many_robot_crossing_is_decided_by_rounding():
    """Ten robots on a 12 m circle all plan through the centre: a symmetric bifurcation (each robot passes left or right)
    that the reference's algorithm resolves by round-off.  Fused vs rounded matrix products then give different
    TRAJECTORIES (metres apart after a few ticks), so for such scenarios `within 1e-9 of the reference` is only
    meaningful against an implementation with the same rounding - which is what the oracle restatement is for.  The test
    records the fact; it is not a bound."""
    worst, _ = _run_pair(scenarios.circle(10, circle_radius=12.0), 5)
    print(f"fma sensitivity symmetric crossing: {worst:.3g}")
    assert worst > 1e-3
