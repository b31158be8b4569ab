"""Per-particle map patches on the CPU oracle (north_star's contact-height variant of EmbodiedSlamFilter::processMap,
src/EmbodiedSlamFilter.cpp:209-227; cloneMaps src/PoseEstimator.cpp:31-45): written inside update(), looked up before
the shared grid, carried through resampling with their particle."""
import numpy as np
import pytest

from slam_eslam_b200 import abi, synthetic

P = 16


def make(oracle, policy, **cfg):
    f = oracle.OraclePoseEstimator(abi.default_config(patch_size=P, **cfg), policy=policy)
    contacts, step = synthetic.setup_filter(f, "C1")
    return f, contacts, step


@pytest.mark.parametrize("policy", [0, 1])
def test_patches_are_written_and_travel_with_resampling(oracle, policy):
    f, contacts, step = make(oracle, policy, min_effective=251)
    cells, origin = f.getPatches(P)
    assert np.isnan(cells[..., 0]).all() and (origin == np.iinfo(np.int32).min).all()
    f.project(step)
    r = f.update(contacts)
    assert r.resampled == 1
    anc = f.getAncestors()
    cells, origin = f.getPatches(P)
    written = ~np.isnan(cells[..., 0])
    assert written.any() and written.reshape(250, -1).sum(axis=1).max() <= 4       # at most one cell per fused contact
    assert (origin != np.iinfo(np.int32).min).all()
    # copies of the same ancestor carry identical maps (cloneMaps: every particle owns its copy)
    for k in range(1, 250):
        if anc[k] == anc[k - 1]:
            assert np.array_equal(cells[k], cells[k - 1], equal_nan=True) and np.array_equal(origin[k], origin[k - 1])
    # written stdev is sqrt(zvar) >= measurement error, heights are near the terrain
    sd = cells[..., 1][written]
    assert (sd >= 0.1 - 1e-6).all() and np.abs(cells[..., 0][written]).max() < 2.0


def test_patch_cells_take_precedence_over_the_shared_grid(oracle):
    """Second update: lookups that land on a written patch cell use it (newest grid first), so results differ from
    the same filter without patches exactly when patch cells were hit."""
    a, contacts, step = make(oracle, 1, min_effective=0)
    b = oracle.OraclePoseEstimator(abi.default_config(min_effective=0), policy=1)
    synthetic.setup_filter(b, "C1")
    still = abi.make_step()                                     # no motion: the same cells are hit again
    for f in (a, b):
        f.project(step)
        f.update(contacts)
    pa, pb = a.getParticles(), b.getParticles()
    assert np.array_equal(pa["weight"], pb["weight"])           # first update: patches were empty
    for f in (a, b):
        f.project(still)
        f.update(contacts)
    pa, pb = a.getParticles(), b.getParticles()
    assert not np.array_equal(pa["weight"], pb["weight"])


def test_literal_and_shared_patches_agree(oracle):
    fa, contacts, step = make(oracle, 0, min_effective=251)
    fb, _, _ = make(oracle, 1, min_effective=251)
    for _ in range(6):
        for f in (fa, fb):
            f.project(step)
            f.update(contacts)
        assert np.array_equal(fa.getAncestors(), fb.getAncestors())
    ca, oa = fa.getPatches(P)
    cb, ob = fb.getPatches(P)
    assert np.array_equal(oa, ob)
    assert np.array_equal(np.isnan(ca), np.isnan(cb))
    np.testing.assert_allclose(np.nan_to_num(ca), np.nan_to_num(cb), rtol=1e-5, atol=1e-6)
