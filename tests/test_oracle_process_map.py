"""EmbodiedSlamFilter::processMap proper on the CPU oracle (src/EmbodiedSlamFilter.cpp:179-232): the scan-grid merge into every
particle's own map and the `match` visual weight p.weight *= pow(w, 0.1).  envire's MLSGrid::merge / match are not vendored
(parity unpinned): these tests pin the DEFINITION the oracle and the CUDA path share (include/eslam_gpu.h
eslam_gpu_process_map), property by property."""
import numpy as np
import pytest

from slam_eslam_b200 import abi

P = 16
INT_MIN = np.iinfo(np.int32).min


def flat_filter(oracle, policy, n=8, height=0.0, stdev=0.05, zsigma=0.1, theta=0.0, **cfg):
    f = oracle.OraclePoseEstimator(abi.default_config(patch_size=P, **cfg), policy=policy)
    cells = np.zeros((200, 200, 2), dtype=np.float32)
    cells[..., 0] = height; cells[..., 1] = stdev
    f.setEnvironment(cells, (0.05, 0.05), (-5.0, -5.0))
    z = np.zeros(n)
    f.setParticles(x=np.linspace(-1, 1, n), y=z + 0.3, theta=z + theta, zpos=z + 0.2, zsigma=z + zsigma, weight=np.full(n, 1.0 / n))
    return f


def scan_grid(w=6, h=4, mean=-0.2, stdev=0.02):
    s = np.zeros((h, w, 2), dtype=np.float32)
    s[..., 0] = mean; s[..., 1] = stdev
    return s, (0.05, 0.05), (0.1, -0.1)            # a 6 x 4 cell scan 0.1 m in front of the body (inside a 16 x 16 cell map)


@pytest.mark.parametrize("policy", [0, 1])
def test_merge_writes_offset_heights_where_the_scan_falls(oracle, policy):
    f = flat_filter(oracle, policy)
    scan, scale, offset = scan_grid()
    f.processMap(scan, scale, offset, match=False, update=True)
    cells, origin = f.getPatches(P)
    p = f.getParticles()
    assert (origin != INT_MIN).all()
    for i in range(len(p["x"])):
        written = ~np.isnan(cells[i, ..., 0])
        assert 20 <= written.sum() <= 6 * 4                           # (nearly) every scan cell lands in its own map cell (theta = 0, same resolution)
        ly, lx = np.nonzero(written)
        # the scan's first cell centre is 0.125 m in front of the particle
        assert lx.min() + origin[i, 0] == int(np.floor((p["x"][i] + 0.1 + 0.025 + 5.0) / 0.05))
        np.testing.assert_allclose(cells[i, ..., 0][written], -0.2 + 0.2, atol=1e-6)      # mean + zPos   (:212 offsetPatch)
        np.testing.assert_allclose(cells[i, ..., 1][written], np.sqrt(0.02 ** 2 + 0.1 ** 2), rtol=1e-6)   # sqrt(stdev^2 + zSigma^2)
    w0 = p["weight"].copy()
    assert np.array_equal(f.getParticles()["weight"], w0)               # update only: weights untouched


def test_merge_of_the_same_scan_twice_is_a_kalman_merge(oracle):
    f = flat_filter(oracle, 1)
    scan, scale, offset = scan_grid()
    f.processMap(scan, scale, offset, match=False, update=True)
    a, _ = f.getPatches(P)
    f.processMap(scan, scale, offset, match=False, update=True)
    b, _ = f.getPatches(P)
    w = ~np.isnan(a[..., 0])
    np.testing.assert_allclose(b[..., 0][w], a[..., 0][w], atol=1e-6)                     # same height again
    np.testing.assert_allclose(b[..., 1][w], a[..., 1][w] / np.sqrt(2.0), rtol=1e-5)      # variance halves


@pytest.mark.parametrize("policy", [0, 1])
def test_match_weight(oracle, policy):
    """A scan that agrees with the map leaves the weights alone (pow(1, 0.1) == 1); one that is d off scales every weight by
    exp(-d^2 / (2 0.2^2)) ^ 0.1; a scan that misses the map everywhere (3-sigma gate) is no information: weight 1."""
    for d, want in ((0.0, 1.0), (0.1, np.exp(-0.01 / 0.08) ** 0.1), (5.0, 1.0)):
        f = flat_filter(oracle, policy, height=0.0)
        scan, scale, offset = scan_grid(mean=-0.2 + d)                  # + zPos 0.2 = d above the map's 0.0
        w0 = f.getParticles()["weight"].copy()
        f.processMap(scan, scale, offset, match=True, update=False)
        np.testing.assert_allclose(f.getParticles()["weight"], w0 * want, rtol=1e-6)      # the match weight is a float upstream
        cells, origin = f.getPatches(P)
        assert np.isnan(cells[..., 0]).all() and (origin == INT_MIN).all()               # match only: maps untouched


def test_match_uses_the_particles_own_map_first(oracle):
    f = flat_filter(oracle, 1)
    scan, scale, offset = scan_grid(mean=-0.1)                           # + zPos = 0.1 above the shared map's 0.0
    f.processMap(scan, scale, offset, match=False, update=True)          # ... now written into every particle's own map
    w0 = f.getParticles()["weight"].copy()
    f.processMap(scan, scale, offset, match=True, update=False)
    np.testing.assert_allclose(f.getParticles()["weight"], w0, rtol=1e-6)                # the own map agrees with the scan: weight 1


def test_rotated_scan_and_policies_agree(oracle):
    rng = np.random.default_rng(2)
    scan = np.zeros((9, 11, 2), dtype=np.float32)
    scan[..., 0] = rng.normal(-0.2, 0.02, scan.shape[:2]); scan[..., 1] = 0.01 + 0.02 * rng.random(scan.shape[:2])
    scan[rng.random(scan.shape[:2]) < 0.2, 0] = np.nan
    out = []
    for policy in (0, 1):
        f = flat_filter(oracle, policy, n=32, theta=0.7)
        f.processMap(scan, (0.04, 0.06), (0.2, -0.3), match=True, update=True)
        f.processMap(scan, (0.04, 0.06), (0.25, -0.3), match=True, update=True)
        out.append((f.getParticles(), f.getPatches(P)))
    (pa, (ca, oa)), (pb, (cb, ob)) = out
    assert np.array_equal(oa, ob)
    assert np.array_equal(np.isnan(ca[..., 0]), np.isnan(cb[..., 0]))
    np.testing.assert_allclose(ca[~np.isnan(ca)], cb[~np.isnan(cb)], rtol=1e-6)
    np.testing.assert_allclose(pa["weight"], pb["weight"], rtol=1e-9)
    assert (~np.isnan(ca[..., 0])).reshape(32, -1).sum(axis=1).min() > 20        # fewer written map cells than the ~80 scan surfaces: several merge per cell


def test_errors(oracle):
    f = oracle.OraclePoseEstimator(abi.default_config(), policy=1)       # no per-particle maps
    z = np.zeros(4)
    f.setParticles(x=z, y=z, theta=z, zpos=z, zsigma=z + 1, weight=z + 0.25)
    scan, scale, offset = scan_grid()
    with pytest.raises(RuntimeError):
        f.processMap(scan, scale, offset)


def test_pose_moments(oracle):
    rng = np.random.default_rng(1)
    n = 3000
    x, y = rng.normal(1.0, 0.5, n), rng.normal(-2.0, 0.2, n)
    w = rng.random(n)
    for policy in (0, 1):
        f = oracle.OraclePoseEstimator(abi.default_config(), policy=policy)
        z = np.zeros(n)
        f.setParticles(x=x, y=y, theta=z, zpos=z, zsigma=z + 1, weight=w)
        m = f.poseMoments()
        W = w.sum()
        mean = np.array([(w * x).sum(), (w * y).sum()]) / W
        d = np.stack([x - mean[0], y - mean[1]])
        cov = (w * d[:, None, :] * d[None, :, :]).sum(axis=2) / W
        np.testing.assert_allclose(m["weight"], W, rtol=1e-13)
        np.testing.assert_allclose(m["mean"], mean, rtol=1e-12)
        np.testing.assert_allclose(m["cov"], cov, rtol=1e-10)
