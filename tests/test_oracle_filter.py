"""The CPU oracle as a filter: properties the reference's tests pin (test/UnitTest.cpp:89-119: resampling keeps N)
plus LITERAL-vs-SHARED agreement, i.e. that the arithmetic the CUDA kernels reproduce bit-for-bit stays within
north_star's tolerances of the line-by-line restatement (relative 1e-5 on normalised weights, 1e-6 m / rad)."""
import numpy as np
import pytest

from slam_eslam_b200 import abi, synthetic


def make_pair(oracle, workload="C1", n=None, **cfg):
    fs = []
    for policy in (oracle.LITERAL, oracle.SHARED):
        f = oracle.OraclePoseEstimator(abi.default_config(**cfg), policy=policy)
        contacts, step = synthetic.setup_filter(f, workload, n=n)
        fs.append(f)
    return fs[0], fs[1], contacts, step


def test_init_matches_and_weights_start_at_zero(oracle):
    lit, sh, _, _ = make_pair(oracle)
    a, b = lit.getParticles(), sh.getParticles()
    assert a["x"].shape == (250,)
    assert np.all(a["weight"] == 0)                        # quirk Q9: the ctor leaves weight = 0
    for k in ("x", "y", "theta"):
        np.testing.assert_allclose(a[k], b[k], rtol=0, atol=1e-14)
    assert abs(a["x"].std() - 0.1) < 0.02 and abs(a["theta"].std() - 0.1) < 0.02
    assert lit.normalizeWeights() == pytest.approx(250.0)  # quirk Q8: sum <= 0 -> uniform 1/N, ESS = N


def test_literal_vs_shared_trajectory_c1(oracle):
    """30 steps of project + update (resampling whenever ESS < 50) on configuration C1."""
    lit, sh, contacts, step = make_pair(oracle)
    resampled = 0
    for i in range(30):
        lit.project(step)
        sh.project(step)
        ra, rb = lit.update(contacts), sh.update(contacts)
        assert ra.resampled == rb.resampled, "step %d" % i
        assert ra.data_particles == rb.data_particles and ra.total_points == rb.total_points
        resampled += ra.resampled
        assert rb.ess == pytest.approx(ra.ess, rel=1e-8)
        assert rb.max_weight == pytest.approx(ra.max_weight, rel=1e-8)
        if ra.resampled:
            assert np.array_equal(lit.getAncestors(), sh.getAncestors()), "step %d" % i
        a, b = lit.getParticles(), sh.getParticles()
        np.testing.assert_allclose(b["weight"], a["weight"], rtol=1e-5, atol=0)      # north_star tolerance
        np.testing.assert_allclose(b["weight"], a["weight"], rtol=1e-8, atol=1e-300)  # what we actually get
        for k in ("x", "y", "theta", "zpos"):
            np.testing.assert_allclose(b[k], a[k], rtol=0, atol=1e-6)
            np.testing.assert_allclose(b[k], a[k], rtol=0, atol=1e-10)
        assert np.array_equal(a["n_found"], b["n_found"]) and np.array_equal(a["floating"], b["floating"])
    assert resampled > 0, "the run never resampled; the test does not exercise the resampling path"


def test_resampling_preserves_particle_count_and_order(oracle):
    """test/UnitTest.cpp:113-117: project / update / resample keeps the particle count."""
    f = oracle.OraclePoseEstimator(abi.default_config(), policy=oracle.LITERAL)
    contacts, step = synthetic.setup_filter(f, "C1")
    for _ in range(10):
        f.project(step)
        f.update(contacts)
        before = f.getParticles()
        f.resample()
        assert f.size() == 250
        anc = f.getAncestors()
        assert np.all(np.diff(anc.astype(np.int64)) >= 0)            # ancestors never decrease (a14)
        after = f.getParticles()
        for k in ("x", "weight", "zpos"):
            assert np.array_equal(after[k], before[k][anc])          # copies keep the ancestor's weight (Q3)
        # stratified: particle i is drawn floor(N w) .. ceil(N w)+1 times
        counts = np.bincount(anc, minlength=250)
        nw = before["weight"] * 250
        assert np.all(counts <= np.ceil(nw) + 1) and np.all(counts >= np.floor(nw) - 1)


def test_cdf_tap_is_the_sequential_sum(oracle):
    f = oracle.OraclePoseEstimator(abi.default_config(), policy=oracle.SHARED)
    rng = np.random.default_rng(0)
    w = rng.random(1000) ** 8
    w /= w.sum()
    z = np.zeros(1000)
    f.setParticles(x=z, y=z, theta=z, zpos=z, zsigma=z + 1, weight=w)
    f.resample()
    cdf = f.getCdf()
    s, ref = 0.0, []
    for v in w:
        s += v
        ref.append(s)
    assert np.array_equal(cdf, np.array(ref))
    anc = f.getAncestors()
    assert np.all(np.diff(anc.astype(np.int64)) >= 0)


def test_systematic_mode_uses_one_uniform(oracle):
    f = oracle.OraclePoseEstimator(abi.default_config(resample_mode=abi.RESAMPLE_SYSTEMATIC), policy=oracle.LITERAL)
    n = 64
    z = np.zeros(n)
    f.setParticles(x=np.arange(n, dtype=float), y=z, theta=z, zpos=z, zsigma=z + 1, weight=np.full(n, 1.0 / n))
    f.resample()
    assert np.array_equal(f.getAncestors(), np.arange(n))            # uniform weights + one offset -> identity


def test_unsigned_exponent_quirk_Q2(oracle):
    """More than 4 fused contacts wraps 4 - cpoints.size() (src/PoseEstimator.cpp:336): every data particle's weight
    becomes 0, the sum is 0 and normalizeWeights falls back to uniform weights."""
    f = oracle.OraclePoseEstimator(abi.default_config(min_effective=0), policy=oracle.LITERAL)
    cells = np.zeros((4, 4, 2), dtype=np.float32)
    cells[..., 1] = 0.05
    f.setEnvironment(cells, (10.0, 10.0), (-20.0, -20.0))
    f.init(16, (0, 0, 0), (0.1, 0.1, 0.1), 0.0, 1.001)
    pts = [(0.1 * i, 0.0, 0.0) for i in range(5)]
    contacts = abi.make_contacts(pts, [0.5] * 5, [-1] * 5)             # 5 ungrouped points -> 5 contacts
    f.project(abi.make_step())
    res = f.update(contacts)
    p = f.getParticles()
    assert np.all(p["n_found"] == 5)
    assert res.weights_sum == 0.0
    assert np.all(p["weight"] == 1.0 / 16)


def test_centroid_and_best_index(oracle):
    for policy in (oracle.LITERAL, oracle.SHARED):
        f = oracle.OraclePoseEstimator(abi.default_config(), policy=policy)
        n = 100
        rng = np.random.default_rng(3)
        x, y, th, z = rng.normal(size=(4, n))
        w = rng.random(n)
        w[37] = 2.0
        f.setParticles(x=x, y=y, theta=th, zpos=z, zsigma=np.ones(n), weight=w)
        assert f.getBestParticleIndex() == 37
        c = f.getCentroid()
        wn = w / w.sum()
        np.testing.assert_allclose(c, [(x * wn).sum(), (y * wn).sum(), (th * wn).sum(), (z * wn).sum()], rtol=1e-12)
        np.testing.assert_allclose(f.getParticles()["weight"], wn, rtol=1e-14)   # getCentroid re-normalises (:356)


def test_resample_variants_with_a_sample_count(oracle):
    """src/ParticleFilter.hpp:85-147: resample_stratified(samples) keeps the ancestors' weights (Q3) and changes the set size;
    resample_multinomial(samples) resets the weights to 1 / old size; both pick particles in proportion to their weight."""
    n = 400
    w = np.zeros(n); w[100] = 0.75; w[300] = 0.25
    z = np.zeros(n)
    for policy in (oracle.LITERAL, oracle.SHARED):
        for multinomial in (False, True):
            f = oracle.OraclePoseEstimator(abi.default_config(), policy=policy)
            f.setParticles(x=np.arange(n, dtype=float), y=z, theta=z, zpos=z, zsigma=z + 1, weight=w)
            (f.resampleMultinomial if multinomial else f.resampleStratified)(1000)
            assert f.size() == 1000
            anc = f.getAncestors()
            assert set(anc.tolist()) == {100, 300}
            frac = (anc == 100).mean()
            assert abs(frac - 0.75) < (0.06 if multinomial else 0.002)
            wt = f.getParticles()["weight"]
            assert np.all(wt == 1.0 / n) if multinomial else np.array_equal(wt, w[anc])
