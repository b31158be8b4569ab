"""Accuracy of the shared host/device arithmetic (slam-eslam_b200/csrc/eslam_math.h), CPU build, against mpmath.
The CUDA build of the same header is compared bit-for-bit with this CPU build in test_gpu_parity.py."""
import math

import mpmath as mp
import numpy as np
import pytest

mp.mp.dps = 40
EXP, LOG, SIN, COS, SIN2PI, COS2PI, RSQRT, H, RATIO_SHARED, RATIO_LITERAL, ROOT_N, POW_U64, BM0, BM1 = range(14)


def h_exact(t):
    t = mp.mpf(t)
    if t < 0:
        x = -t
        return 1 / (mp.sqrt(mp.pi / 2) * mp.exp(x * x / 2) * mp.erfc(x / mp.sqrt(2)))
    return mp.npdf(t) / mp.ncdf(t)


def max_rel(oracle, fn, ref, xs, b=0.0):
    worst = 0.0
    for x in xs:
        r = ref(x)
        if r == 0:
            continue
        worst = max(worst, float(abs((mp.mpf(oracle.math_fn(fn, x, b)) - r) / r)))
    return worst


def test_exp_log(oracle):
    rng = np.random.default_rng(1)
    assert max_rel(oracle, EXP, lambda x: mp.exp(mp.mpf(x)), rng.uniform(-700, 700, 500)) < 3e-16
    assert max_rel(oracle, EXP, lambda x: mp.exp(mp.mpf(x)), rng.uniform(-40, 0, 500)) < 3e-16
    assert oracle.math_fn(EXP, 0.0) == 1.0
    assert oracle.math_fn(EXP, -1000.0) == 0.0
    assert math.isinf(oracle.math_fn(EXP, 800.0))
    assert max_rel(oracle, LOG, lambda x: mp.log(mp.mpf(x)), rng.uniform(0, 1, 500)) < 5e-16
    assert max_rel(oracle, LOG, lambda x: mp.log(mp.mpf(x)), 10 ** rng.uniform(-300, 300, 500)) < 5e-16
    assert oracle.math_fn(LOG, 1.0) == 0.0


def test_sincos(oracle):
    rng = np.random.default_rng(2)
    for x in rng.uniform(-200, 200, 500):
        assert abs(oracle.math_fn(SIN, x) - float(mp.sin(mp.mpf(x)))) < 3e-16
        assert abs(oracle.math_fn(COS, x) - float(mp.cos(mp.mpf(x)))) < 3e-16
    for u in rng.uniform(0, 1, 500):
        assert abs(oracle.math_fn(SIN2PI, u) - float(mp.sin(2 * mp.pi * mp.mpf(u)))) < 3e-16
        assert abs(oracle.math_fn(COS2PI, u) - float(mp.cos(2 * mp.pi * mp.mpf(u)))) < 3e-16
    assert oracle.math_fn(SIN2PI, 0.25) == 1.0 and oracle.math_fn(COS2PI, 0.5) == -1.0


def test_rsqrt(oracle):
    rng = np.random.default_rng(3)
    assert max_rel(oracle, RSQRT, lambda x: 1 / mp.sqrt(mp.mpf(x)), 10 ** rng.uniform(-29, 29, 1000)) < 4e-16
    assert max_rel(oracle, RSQRT, lambda x: 1 / mp.sqrt(mp.mpf(x)), 10 ** rng.uniform(-200, 200, 200)) < 4e-16


@pytest.mark.parametrize("lo,hi,tol", [(-40, -10, 1e-14), (-10, 0, 3e-12), (0, 4, 3e-10), (4, 7, 1.5e-9),
                                       (7, 10, 3e-9), (10, 30, 2e-13)])
def test_inverse_mills_ratio(oracle, lo, hi, tol):
    """h(t) = pdf/cdf of the unit normal: the reference's contactLikelihoodRatio core (src/ContactModel.cpp:104-115).
    Inside the table (|t| < 10) the bound is set by the fp32 evaluation of the z^2 q(z) part (eslam_math.h h_poly)."""
    rng = np.random.default_rng(4)
    assert max_rel(oracle, H, h_exact, rng.uniform(lo, hi, 400)) < tol


def test_likelihood_ratio_shared_vs_literal(oracle):
    """Table-driven ratio against the boost-style pdf/cdf through libm exp/erfc, inside the 3-sigma gate
    (|z| < 3 sqrt(zvar), src/PoseEstimator.hpp:101), wherever the ratio can reach the 1e-9 group threshold."""
    rng = np.random.default_rng(5)
    worst = 0.0
    for _ in range(4000):
        zvar = 10 ** rng.uniform(-2.2, 0.3)
        z = rng.uniform(-3, 3) * math.sqrt(zvar)
        a, b = oracle.math_fn(RATIO_SHARED, z, zvar), oracle.math_fn(RATIO_LITERAL, z, zvar)
        if b > 1e-10:
            worst = max(worst, abs(a - b) / b)
    assert worst < 3e-9


def test_powers(oracle):
    rng = np.random.default_rng(6)
    for w in rng.uniform(1e-12, 1, 300):
        for n in (1, 2, 3, 4, 5, 7):
            assert abs(oracle.math_fn(ROOT_N, w, n) - w ** (1.0 / n)) <= 4e-16 * w ** (1.0 / n)
    assert oracle.math_fn(POW_U64, 0.81, 0) == 1.0
    assert oracle.math_fn(POW_U64, 0.5, 3) == 0.125
    assert oracle.math_fn(POW_U64, 0.9, 2 ** 63) == 0.0          # quirk Q2: wrapped size_t exponent
    assert oracle.math_fn(POW_U64, 1.0, 2 ** 63) == 1.0


def test_cube_root(oracle):
    """pow(weight, 1/3) (three of four contact groups found, src/PoseEstimator.cpp:309) as a division-free cube root."""
    rng = np.random.default_rng(8)
    xs = np.concatenate([10 ** rng.uniform(-300, 300, 1500), rng.uniform(1e-12, 1, 1500), [1e-310, 5e-324, 1.7e307]])
    assert max_rel(oracle, ROOT_N, lambda x: mp.cbrt(mp.mpf(x)), xs, 3) < 1.2 * 2.0 ** -53
    for k in (1.0, 2.0, 3.0, 0.5, 1e5):
        assert oracle.math_fn(ROOT_N, k ** 3, 3) == k                     # exact cubes come out exact
    assert oracle.math_fn(ROOT_N, 0.0, 3) == 0.0
    assert math.isinf(oracle.math_fn(ROOT_N, float("inf"), 3))


def test_box_muller_moments(oracle):
    rng = np.random.default_rng(7)
    u = rng.uniform(1e-12, 1, (20000, 2))
    n0 = np.array([oracle.math_fn(BM0, a, b) for a, b in u])
    n1 = np.array([oracle.math_fn(BM1, a, b) for a, b in u])
    for n in (n0, n1):
        assert abs(n.mean()) < 0.03 and abs(n.std() - 1) < 0.03
    assert abs(np.corrcoef(n0, n1)[0, 1]) < 0.03
