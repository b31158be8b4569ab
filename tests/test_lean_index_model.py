"""The fused fixed-point cell index of the lean weight kernel (lean_eval in slam-eslam_b200/csrc/eslam_kernels.cuh, constants
from fill_fast_params in eslam_gpu.cu), restated with exact-rational fma emulation and checked on the CPU:

    per particle  u0 = fma(x, 1/scale, (1.5 2^28 - offset/scale) - tile_x0)
    per contact   u  = fma(c/scale, px, fma(-s/scale, py, u0));   cell = low_word(u) >> 24   (tile-relative)
    the decision is TRUSTED unless the 24-bit fraction lies within 4 / 2^24 of a cell border

against the cell the reference defines, floor((wx - offset)/scale) with wx the transformed contact point (envire toGrid;
SURVEY.md §8 a10: the map-cell index that must be bit-exact).  Claim under test: a trusted decision IS the exact cell; the
untrusted ones (which the kernel recomputes with the IEEE division) are rare."""
import math
import struct
from fractions import Fraction

import numpy as np


def rn(fr):
    """Fraction -> nearest double (ties to even): float(Fraction) is correctly rounded in CPython."""
    return float(fr)


def fma(a, b, c):
    return rn(Fraction(a) * Fraction(b) + Fraction(c))


def low_word(x):
    return struct.unpack("<Q", struct.pack("<d", x))[0] & 0xFFFFFFFF


def test_trusted_lean_cells_are_the_exact_cells():
    rng = np.random.default_rng(11)
    checked = untrusted = 0
    for scale, W, tile_w in ((0.05, 2048, 156), (0.1, 512, 156), (0.037, 4096, 156)):
        inv = 1.0 / scale
        offset = -W * scale / 2.0
        c0 = 402653184.0 - offset * inv                     # fc.c0x: product and difference rounded separately (-fmad=false)
        for _ in range(6000):
            x0 = int(rng.integers(0, (W - tile_w) // 2 + 1)) * 2            # even tile origin inside the grid
            # a particle whose footprint lies inside the tile: own cell at least 13 cells from the tile edge, contacts <= 0.6 m
            cell_p = x0 + rng.uniform(13.0, tile_w - 13.0)
            x = offset + cell_p * scale
            y = rng.uniform(-40.0, 40.0)
            th = rng.uniform(-7.0, 7.0)
            s, c = math.sin(th), math.cos(th)
            r, a = rng.uniform(0.0, min(0.6, 12.0 * scale)), rng.uniform(0, 2 * math.pi)
            px, py = r * math.cos(a), r * math.sin(a)
            # lean path
            u0 = fma(x, inv, c0 - float(x0))
            axx, axy = c * inv, -s * inv
            u = fma(axx, px, fma(axy, py, u0))
            k = low_word(u)
            cell_lean = k >> 24
            trusted = ((k * 256 + 1024) & 0xFFFFFFFF) >= 2048
            # exact path: the SHARED transform, then the division as written
            wx = fma(c, px, fma(-s, py, x))
            cell_exact = math.floor((wx - offset) / scale) - x0
            checked += 1
            if not trusted:
                untrusted += 1
                continue
            assert cell_lean == cell_exact, (scale, x0, x, th, px, py, cell_lean, cell_exact)
    assert checked == 18000 and untrusted < 40               # ~ 8 / 2^24 * 2 axes per lookup would be 1e-6; borders drawn
    #                                                          on purpose below make up the rest of the budget


def test_points_on_cell_borders_are_untrusted_or_exact():
    """Contact points placed (to within a few ulps) ON cell borders: the lean decision must either decline or be right."""
    rng = np.random.default_rng(12)
    scale, W, tile_w = 0.05, 2048, 156
    inv, offset = 1.0 / scale, -W * scale / 2.0
    c0 = 402653184.0 - offset * inv
    declined = 0
    for _ in range(4000):
        x0 = int(rng.integers(0, (W - tile_w) // 2 + 1)) * 2
        border = x0 + int(rng.integers(20, tile_w - 20))
        px, py = rng.uniform(-0.4, 0.4), rng.uniform(-0.4, 0.4)
        th = rng.uniform(-3.0, 3.0)
        s, c = math.sin(th), math.cos(th)
        # choose x so that the transformed point lands on the border, then nudge it by up to +-3 ulps
        x = (offset + border * scale) - (c * px - s * py)
        x = float(np.nextafter(x, x + rng.choice([-1.0, 1.0]))) if rng.random() < 0.7 else x
        u = fma(c * inv, px, fma(-s * inv, py, fma(x, inv, c0 - float(x0))))
        k = low_word(u)
        trusted = ((k * 256 + 1024) & 0xFFFFFFFF) >= 2048
        wx = fma(c, px, fma(-s, py, x))
        cell_exact = math.floor((wx - offset) / scale) - x0
        if trusted:
            assert (k >> 24) == cell_exact
        else:
            declined += 1
    assert declined > 3900                                    # on a border the kernel hands the decision to the IEEE division


def test_trusted_lean_gate_decisions_are_the_exact_gate(oracle):
    """The 3-sigma patch gate of MLSGrid::get, d^2 < 9 zvar, as the lean kernel takes it: on the HIGH WORD of |t|,
    t = d * (rsqrt(zvar) / corr), against 3 / corr; a decision within one high-word step of the threshold is not trusted
    (fill_fast_params: gate_lo = hi(3/corr) - 1, gate_span = 2).  Trusted decisions must equal the gate as written."""
    RSQRT = 6
    corr = 0.33
    inv_corr = 1.0 / corr
    T = 3.0 * inv_corr

    def hi(x):
        return struct.unpack("<Q", struct.pack("<d", x))[0] >> 32
    gate_lo, gate_span = hi(T) - 1, 2
    rng = np.random.default_rng(13)
    untrusted = flips_if_trusted_blindly = 0
    n = 40000
    for i in range(n):
        zvar = float(10.0 ** rng.uniform(-4, 1))
        if i % 2:
            d = 3.0 * math.sqrt(zvar) * (1.0 + rng.choice([-1.0, 1.0]) * 10.0 ** rng.uniform(-16, -3))   # around the threshold
        else:
            d = float(rng.normal(0.0, 4.0 * math.sqrt(zvar)))
        if rng.random() < 0.5:
            d = -d
        qc = oracle.math_fn(RSQRT, zvar) * inv_corr
        t = d * qc
        dlt = ((hi(t) & 0x7FFFFFFF) - gate_lo) & 0xFFFFFFFF
        inside = dlt >= 0x80000000
        trusted = dlt > gate_span
        exact = d * d < 9.0 * zvar
        if not trusted:
            untrusted += 1
            flips_if_trusted_blindly += int(inside != exact)
            continue
        assert inside == exact, (d, zvar, t, dlt)
    assert 0 < untrusted < n // 2          # only the draws placed within ~1e-6 of the threshold are declined
