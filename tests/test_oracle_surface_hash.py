"""SurfaceHash, the slope-indexed relocaliser (src/SurfaceHash.hpp:113-232; PoseEstimator::init(n, hash) src/PoseEstimator.cpp:75-86;
sampleFromHash :130-182), on the CPU oracle: the definition the CUDA path shares (candidate order, skipped footprint points,
Philox draws instead of libc rand(): include/eslam_gpu.h eslam_gpu_hash_create)."""
import numpy as np
import pytest

from slam_eslam_b200 import abi, synthetic


def flat(oracle, height=0.3, n=60):
    f = oracle.OraclePoseEstimator(abi.default_config(), policy=oracle.SHARED)
    cells = np.zeros((n, n, 2), dtype=np.float32)
    cells[..., 0] = height; cells[..., 1] = 0.02
    f.setEnvironment(cells, (0.05, 0.05), (-n * 0.025, -n * 0.025))
    return f


def test_flat_map_has_one_bucket_and_all_interior_cells(oracle):
    f = flat(oracle)
    f.hashCreate(20, 16)
    n, counts = f.hashInfo()
    assert counts.sum() == n
    centre = [bx * 20 + by for bx in (9, 10) for by in (9, 10)]    # slope (0, 0) sits on the edge between buckets 9 and 10: rounding decides
    assert counts[centre].sum() == n
    assert 16 * 40 * 40 < n <= 16 * 60 * 60                # the footprint (0.5 m) must see >= 3 cells: a border strip drops out
    c = f.hashGet(0, n)
    np.testing.assert_allclose(c["z"], 0.3 + 0.18, atol=1e-6)                  # pose.z() + 0.18 (:216)
    assert set(np.round(c["theta"] / (2 * np.pi / 16)).astype(int)) == set(range(16))
    assert np.all(np.diff(c["theta"]) >= 0)                # creation order: heading first


def test_tilted_plane_slopes_rotate_with_the_heading(oracle):
    """z = 0.2 x: the plane under the footprint has gradient (0.2, 0) in the world; in the body frame of a footprint rotated by
    phi = (a + 1) 2 pi / 16 (the reference rotates before use, :177-181) the slopes are (0.2 cos phi, -0.2 sin phi)."""
    n = 80
    f = oracle.OraclePoseEstimator(abi.default_config(), policy=oracle.SHARED)
    xs = -n * 0.025 + (np.arange(n) + 0.5) * 0.05
    cells = np.zeros((n, n, 2), dtype=np.float32)
    cells[..., 0] = (0.2 * xs)[None, :]; cells[..., 1] = 0.02
    f.setEnvironment(cells, (0.05, 0.05), (-n * 0.025, -n * 0.025))
    f.hashCreate(20, 16)
    total, _ = f.hashInfo()
    c = f.hashGet(0, total)
    a = np.round(c["theta"] / (2 * np.pi / 16)).astype(int)
    phi = (a + 1) * 2 * np.pi / 16
    bx = np.clip(np.floor((0.2 * np.cos(phi) + 1) / 2 * 20), 0, 19).astype(int)
    by = np.clip(np.floor((-0.2 * np.sin(phi) + 1) / 2 * 20), 0, 19).astype(int)
    clear = (np.abs(np.cos(phi)) > 1e-6) & (np.abs(np.sin(phi)) > 1e-6)     # a zero slope component sits on the edge between buckets 9 and 10
    agree = (c["bucket"] == bx * 20 + by)[clear].mean()
    assert agree > 0.98 and clear.mean() > 0.7             # cell quantisation moves a few candidates next to a bucket edge


def test_init_from_hash_and_errors(oracle):
    f = flat(oracle)
    with pytest.raises(RuntimeError):
        f.initFromHash(10)                                 # "could not sample from pose hash." (src/PoseEstimator.cpp:84)
    f.hashCreate(20, 16)
    f.initFromHash(500)
    p = f.getParticles()
    assert f.size() == 500 and np.all(p["weight"] == 0) and np.all(p["zsigma"] == 0) and np.all(p["floating"] == 1)
    np.testing.assert_allclose(p["zpos"], 0.48, atol=1e-6)
    cx = (p["x"] + 60 * 0.025) / 0.05 - 0.5
    assert np.abs(cx - np.round(cx)).max() < 1e-9          # candidate poses sit on cell centres
    assert len(np.unique(np.round(p["theta"], 6))) > 8


def test_sample_from_hash_replaces_the_lowest_weights(oracle):
    f = oracle.OraclePoseEstimator(abi.default_config(), policy=oracle.SHARED)
    contacts, step = synthetic.setup_filter(f, "C1", n=2000)
    f.hashCreate(20, 16)
    n_poses, counts = f.hashInfo()
    rng = np.random.default_rng(0)
    w = rng.random(2000) + 0.5
    f.setParticles(weight=w)
    before = f.getParticles()
    replaced = f.sampleFromHash(contacts, replace_percentage=0.05, avg_factor=0.1)
    after = f.getParticles()
    changed = np.nonzero(after["zsigma"] == 0.5)[0]
    assert replaced == len(changed)
    if replaced:                                           # relevance^3 >= 0.8: the bucket of the body's own slopes is rare enough
        assert 0.8 * 100 <= replaced <= 100
        assert set(changed) == set(np.argsort(w.astype(np.float32), kind="stable")[:replaced])   # the lowest weights (:147-153)
        assert len(np.unique(after["weight"][changed])) == 1                 # getWeightAvg() * avgFactor * relevance^3 (:162)
        assert 0.8 * w.mean() * 0.1 <= after["weight"][changed][0] <= w.mean() * 0.1 * (1 + 1e-12)
        assert np.all(after["floating"][changed] == 1)
    keep = np.setdiff1d(np.arange(2000), changed)
    for k in ("x", "y", "theta", "zpos", "zsigma", "weight"):
        assert np.array_equal(after[k][keep], before[k][keep])
    # a map whose candidates (nearly) all share the body's bucket: relevance^3 < 0.8 -> nothing is replaced (:158-159)
    g = oracle.OraclePoseEstimator(abi.default_config(), policy=oracle.SHARED)
    xs = -1.5 + (np.arange(60) + 0.5) * 0.05
    cells = np.zeros((60, 60, 2), dtype=np.float32)
    cells[..., 0] = (0.02 * xs)[None, :] + (0.03 * xs)[:, None]; cells[..., 1] = 0.02          # a gentle slope: bucket (10, 10) for every heading
    g.setEnvironment(cells, (0.05, 0.05), (-1.5, -1.5))
    g.hashCreate(20, 16)
    g.initFromHash(300)
    g.setParticles(weight=np.full(300, 1.0 / 300))
    tilted = abi.make_contacts([(0.25, 0.25, 0.004), (0.25, -0.25, 0.0), (-0.25, 0.25, 0.002), (-0.25, -0.25, -0.002)], [0.5] * 4, [-1] * 4)
    assert g.sampleFromHash(tilted) == 0
