"""The partition rule of k_resample_stream (slam-eslam_b200/csrc/eslam_kernels.cuh), restated in numpy and checked on the CPU
against the oracle's stratified walk: a persistent CTA owns strata [K0, K1); for every SOURCE tile t it visits the candidate
strata floor(n S_t) - 2 .. floor(n S_t+1) + 2 (S_t = exact running sum before the tile) and serves a stratum iff
    (t == 0 or S_t < sum_r) and (t last or not cdf[last of t] < sum_r)
Every stratum must be served by exactly one (CTA, tile), with the ancestor the reference's walk finds -- on balanced weights
and on weights with zero tiles, ties, one heavy particle, a sum below the last draw."""
import numpy as np
import pytest

from slam_eslam_b200 import abi

TILE = 1024


def stream_partition(cdf, u, n_cta):
    n = len(cdf)
    n_tiles = (n + TILE - 1) // TILE
    start = np.r_[0.0, cdf[TILE - 1::TILE][:n_tiles - 1]] if n_tiles > 1 else np.array([0.0])   # S_t == cdf[1024 t - 1]
    per = (((n + n_cta - 1) // n_cta) + 31) & ~31
    anc = np.full(n, -1, dtype=np.int64)
    served = np.zeros(n, dtype=np.int32)
    loads = 0
    dn = float(n)
    for c in range(n_cta):
        K0, K1 = min(n, c * per), min(n, c * per + per)
        if K0 >= K1:
            continue
        c_lo, c_hi = int((start < K0 / dn).sum()), int((start < K1 / dn).sum())
        tA, tB = max(c_lo - 1, 0), max(c_hi - 1, 0)
        for t in range(tA, tB + 1):
            klo, khi = K0, K1
            if t > 0:
                f = np.floor(start[t] * dn) - 2.0
                if f > K0:
                    klo = K1 if f >= K1 else int(f)
            if t + 1 < n_tiles:
                cc = np.floor(start[t + 1] * dn) + 3.0
                if cc < K1:
                    khi = K0 if cc <= K0 else int(cc)
            if klo >= khi:
                continue                                   # the tile is never loaded
            loads += 1
            e0 = t * TILE
            cnt = min(TILE, n - e0)
            tile = cdf[e0:e0 + cnt]
            k = np.arange(klo, khi)
            sum_r = (k.astype(np.float64) + u[k]) / dn
            mine = ((t == 0) | (start[t] < sum_r)) & ((t + 1 == n_tiles) | ~(tile[-1] < sum_r))
            lo = np.minimum(np.searchsorted(tile, sum_r[mine], side="left"), cnt - 1)      # first element with !(cdf < sum_r)
            anc[k[mine]] = e0 + lo
            served[k[mine]] += 1
    return anc, served, loads


def weight_vectors(n, rng):
    w = rng.random(n) + 0.5
    yield "balanced", w
    z = w.copy(); z[2 * TILE:5 * TILE] = 0.0; z[:TILE // 2] = 0.0
    yield "zero tiles", z
    yield "ties", np.repeat(rng.random(n // 8 + 1), 8)[:n]
    h = np.full(n, 1e-12); h[n // 3] = 1.0
    yield "one heavy particle", h
    yield "sum below the last draws", w / w.sum() * 0.999
    yield "steep", rng.random(n) ** 12 + 1e-9


@pytest.mark.parametrize("n,n_cta", [(250, 148), (10000, 148), (40000, 7), (65536, 148)])
def test_every_stratum_is_served_once_with_the_walks_ancestor(oracle, n, n_cta):
    rng = np.random.default_rng(n)
    zeros = np.zeros(n)
    for name, w in weight_vectors(n, rng):
        o = oracle.OraclePoseEstimator(abi.default_config(), policy=oracle.LITERAL)
        o.setParticles(x=zeros, y=zeros, theta=zeros, zpos=zeros, zsigma=zeros + 1, weight=w)
        if name != "sum below the last draws":
            o.normalizeWeights()
        wn = o.getParticles()["weight"]
        o.resample()
        want, cdf = o.getAncestors().astype(np.int64), o.getCdf()
        assert np.array_equal(cdf, np.cumsum(wn))           # the oracle's running sum is the sequential one
        # any uniforms in [0, 1) define a stratified walk over this running sum: draw our own and compare the partition with
        # the plain lower-bound search (what src/ParticleFilter.hpp:96-105 computes, clamped at the last particle)
        u = rng.random(n)
        sum_r = (np.arange(n, dtype=np.float64) + u) / float(n)
        ref = np.minimum(np.searchsorted(cdf, sum_r, side="left"), n - 1)
        anc, served, loads = stream_partition(cdf, u, n_cta)
        assert np.all(served == 1), (name, np.flatnonzero(served != 1)[:5])
        assert np.array_equal(anc, ref), name
        assert np.all(np.diff(want) >= 0) and np.all(np.diff(anc) >= 0)      # both walks are monotone
        n_tiles = (n + TILE - 1) // TILE
        assert loads <= n_tiles + 2 * n_cta                 # every tile once, plus at most the boundary tiles of every CTA
