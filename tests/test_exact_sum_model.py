"""The exact running sum of the resampling walk (DESIGN.md §4 "Exact running sum"; classify_elem / segf_compose / segf_apply in
slam-eslam_b200/csrc/eslam_kernels.cuh), restated in Python and checked on the CPU against the sequential fp64 sum the
reference computes (src/ParticleFilter.hpp:95,102):

  * every element is NEUTRAL (adding it cannot change the sum), INTERIOR (the add is an integer add of rne(w / ulp) on the
    bits of the sum, with a tie case that depends only on the parity of the sum: a 2-state function (d0, d1)), or BOUNDARY
    (added in real fp64) -- decided from an APPROXIMATE prefix only;
  * functions of consecutive interior elements compose associatively.
Applying the classes in order must reproduce the sequential sum bit for bit; composing runs first must give the same."""
import struct

import numpy as np
import pytest

BIN_NONE = -100000


def d2u(x):
    return struct.unpack("<Q", struct.pack("<d", float(x)))[0]


def u2d(u):
    return struct.unpack("<d", struct.pack("<Q", u & 0xFFFFFFFFFFFFFFFF))[0]


def classify(w, a_prev, a_cur, eps):
    """kind 0 neutral / 1 interior (b, d0, d1) / 2 boundary -- classify_elem, statement by statement."""
    if w == 0.0:
        return (0, BIN_NONE, 0, 0)
    if not (w > 0.0) or not (a_cur < 1.7e308):
        return (2, BIN_NONE, 0, 0)
    lo = a_prev - a_prev * eps            # the kernel uses one fma; any value within rounding of a_prev (1 - eps) will do
    hi = a_cur + a_cur * eps
    ulo, uhi, uw = d2u(lo), d2u(hi), d2u(w)
    if not (lo > 0.0) or (ulo >> 52) == 0:
        return (2, BIN_NONE, 0, 0)
    blo, bhi = (ulo >> 52) - 1023, (uhi >> 52) - 1023
    ew = -1023 if (uw >> 52) == 0 else (uw >> 52) - 1023
    if ew + 1 <= blo - 54:
        return (0, BIN_NONE, 0, 0)        # w < 2^(blo - 54): under a quarter ulp of the sum
    if blo != bhi or (uw >> 52) == 0:
        return (2, BIN_NONE, 0, 0)
    k = blo - ew
    if k < 0:
        return (2, BIN_NONE, 0, 0)
    M = (uw & 0x000FFFFFFFFFFFFF) | 0x0010000000000000
    if k == 0:
        return (1, blo, M, M)
    fl, frac, half = M >> k, M & ((1 << k) - 1), 1 << (k - 1)
    up = 1 if frac > half else 0
    tie = frac == half
    d0 = fl + up + (1 if (tie and (fl & 1)) else 0)
    d1 = fl + up + (1 if (tie and not (fl & 1)) else 0)
    return (1, blo, d0, d1)


def apply(S, f):
    """segf_apply: the integer add on the bits of S; returns (S', bad)."""
    b, d0, d1 = f
    if b == BIN_NONE:
        return S, False
    u = d2u(S)
    if (u >> 52) - 1023 != b:
        return S, True
    return u2d(u + (d1 if (u & 1) else d0)), False


def compose(a, b):
    """segf_compose: a then b (b = None binade means identity)."""
    ab, ad0, ad1 = a
    bb, bd0, bd1 = b
    cb = bb if bb != BIN_NONE else ab
    bad = ab != BIN_NONE and bb != BIN_NONE and ab != bb
    c0 = ad0 + (bd1 if (ad0 & 1) else bd0)
    c1 = ad1 + (bd1 if ((1 + ad1) & 1) else bd0)
    return (cb, c0, c1), bad


def approximate_prefix(w):
    """An inclusive prefix with a DIFFERENT rounding than the sequential sum (blocked, like the tile scans of the kernels)."""
    n = len(w)
    blocks = [w[i:i + 1024] for i in range(0, n, 1024)]
    offs = np.r_[0.0, np.cumsum([b.sum() for b in blocks])[:-1]]       # pairwise block sums
    return np.concatenate([o + np.cumsum(b) for o, b in zip(offs, blocks)])


def vectors(n, rng):
    w = rng.random(n); yield "uniform", w / w.sum()
    w = rng.random(n) ** 20; yield "steep", w / w.sum()
    w = np.full(n, 1.0 / n); yield "equal (ties)", w
    w = np.r_[np.zeros(n // 3), rng.random(n - n // 3)]; yield "leading zeros", w / w.sum()
    w = np.full(n, 1e-17); w[5] = 1.0; yield "one heavy + quarter-ulp tail", w
    w = 2.0 ** rng.integers(-60, -8, n).astype(np.float64); yield "powers of two", w / w.sum()
    w = rng.random(n) * 1e3; yield "unnormalised", w


@pytest.mark.parametrize("n", [3000, 20000])
def test_classes_reproduce_the_sequential_sum(n):
    rng = np.random.default_rng(n)
    eps = max(8 * n * 2.0 ** -53, 2.0 ** -40)
    for name, w in vectors(n, rng):
        A = approximate_prefix(w)
        seq = np.cumsum(w)                                   # numpy's cumsum is the left-to-right fp64 sum
        S, kinds = 0.0, [0, 0, 0]
        run = (BIN_NONE, 0, 0)                               # composed function of the current run of interior elements
        S_run_start = 0.0
        for i in range(n):
            kind, b, d0, d1 = classify(float(w[i]), float(A[i - 1]) if i else 0.0, float(A[i]), eps)
            kinds[kind] += 1
            if kind == 2:
                S = S + float(w[i]); run = (BIN_NONE, 0, 0); S_run_start = S
            elif kind == 1:
                S, bad = apply(S, (b, d0, d1))
                assert not bad, (name, i)
                run, bad = compose(run, (b, d0, d1))
                assert not bad, (name, i)
                S2, bad = apply(S_run_start, run)            # the composed run applied to the sum at its start: same bits
                assert not bad and S2 == S, (name, i)
            assert S == seq[i], "%s: element %d (kind %d) %r != %r" % (name, i, kind, S, seq[i])
        assert kinds[2] < 200, (name, kinds)                 # boundaries are rare: first element, binade crossings


def test_composition_is_associative():
    rng = np.random.default_rng(0)
    for _ in range(2000):
        b = int(rng.integers(-30, 0))
        fs = []
        for _ in range(3):
            fl = int(rng.integers(0, 1 << 40))
            d0 = fl + int(rng.integers(0, 2)); d1 = fl + int(rng.integers(0, 2))
            fs.append((b, d0, d1) if rng.random() < 0.9 else (BIN_NONE, 0, 0))
        (l, _), (r, _) = compose(compose(fs[0], fs[1])[0], fs[2]), compose(fs[0], compose(fs[1], fs[2])[0])
        assert l == r
        # and composition means "apply one after the other", for an even and an odd mantissa
        for mant in (0x0010000000000000 | int(rng.integers(0, 1 << 30)) << 1, (0x0010000000000000 | int(rng.integers(0, 1 << 30)) << 1) + 1):
            S0 = u2d(((b + 1023) << 52) | (mant & 0x000FFFFFFFFFFFFF))
            S, ok = S0, True
            for f in fs:
                S2, bad = apply(S, f)
                if bad or (d2u(S2) >> 52) != (d2u(S0) >> 52):
                    ok = False; break                        # left the binade: the kernels make such an element a boundary
                S = S2
            if ok and l[0] != BIN_NONE:
                assert apply(S0, l)[0] == S


def test_branches_of_a_composed_function_differ_by_at_most_one():
    """k_norm_scan stores, per element, d0 of the composed rounding function and d1 - d0 in two bits (k_cdf): two walks that
    start one ulp apart are 0, 1 or 2 ulps apart ever after, so d1 - d0 is -1, 0 or +1 for ANY run of interior elements."""
    rng = np.random.default_rng(1)
    for _ in range(300):
        b = int(rng.integers(-30, 0))
        run = (BIN_NONE, 0, 0)
        for _ in range(int(rng.integers(1, 60))):
            # single-element functions as classify() makes them: non-tie (d0 == d1) or tie (they differ by one, by the parity of fl)
            fl = int(rng.integers(0, 1 << 30))
            if rng.random() < 0.5:
                f = (b, fl, fl)
            else:
                f = (b, fl + (1 if fl & 1 else 0), fl + (0 if fl & 1 else 1))
            run, bad = compose(run, f)
            assert not bad
            assert run[2] - run[1] in (-1, 0, 1), run


@pytest.mark.parametrize("n", [5000])
def test_offsets_from_a_base_reproduce_the_sequential_sum(n):
    """What k_cdf does: running sum after element i = bits(base) + d0 (+ d1 - d0 if the base's mantissa is odd), base = the sum at
    the last boundary element (or at the start)."""
    rng = np.random.default_rng(n)
    eps = max(8 * n * 2.0 ** -53, 2.0 ** -40)
    for name, w in vectors(n, rng):
        A = approximate_prefix(w)
        seq = np.cumsum(w)
        base, run = 0.0, (BIN_NONE, 0, 0)
        for i in range(n):
            kind, b, d0, d1 = classify(float(w[i]), float(A[i - 1]) if i else 0.0, float(A[i]), eps)
            if kind == 2:
                base = (seq[i - 1] if i else 0.0) + float(w[i]); run = (BIN_NONE, 0, 0)
                got = base
            else:
                if kind == 1:
                    run, bad = compose(run, (b, d0, d1))
                    assert not bad
                if run[0] == BIN_NONE:
                    got = base
                else:
                    u = d2u(base)
                    delta = run[2] - run[1]
                    assert delta in (-1, 0, 1)
                    got = u2d(u + run[1] + (delta if (u & 1) else 0))
            assert got == seq[i], (name, i)
