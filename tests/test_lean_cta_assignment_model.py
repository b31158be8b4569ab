"""Where the CTAs of the lean weight kernel start (k_bin_scatter, block 0; DESIGN.md §4 "The lean path"), restated in numpy:
one CTA per non-empty bin first (up to the number of CTAs), the others where sorted position (i + 1/2) M / R lies, i.e. in
proportion to the bins' particles.  Checked here: every non-empty bin of a wide cloud is started on, no CTA starts on an empty
bin, and the crews of the big bins are within one CTA of their proportional share."""
import numpy as np
import pytest


def cta_start_bins(counts, n_cta):
    """k_bin_scatter's assignment, statement by statement."""
    counts = np.asarray(counts, dtype=np.int64)
    offs = np.concatenate([[0], np.cumsum(counts)])          # exclusive scan, offs[nb] = M
    M = int(offs[-1])
    nonempty = np.flatnonzero(counts > 0)
    first = min(len(nonempty), n_cta)
    out = list(nonempty[:first])
    R = n_cta - first
    for i in range(R):
        p = ((2 * i + 1) * M) // (2 * R)
        b = int(np.searchsorted(offs[:-1], p, side="right")) - 1   # last bin whose offset is <= p
        out.append(b)
    return np.array(out, dtype=np.int64)


def wide_cloud_counts(rng, n, sigma_cells, bsx=232, bsy=76, nbx=9, nby=27, centre=(1024, 1024)):
    x = rng.normal(centre[0], sigma_cells, n); y = rng.normal(centre[1], sigma_cells, n)
    b = (np.clip(y, 0, 2047) // bsy).astype(int) * nbx + (np.clip(x, 0, 2047) // bsx).astype(int)
    return np.bincount(b, minlength=nbx * nby)


@pytest.mark.parametrize("sigma", [2.0, 10.0, 40.0, 120.0])
def test_every_nonempty_bin_is_started_on_and_crews_are_proportional(sigma):
    rng = np.random.default_rng(int(sigma))
    n, G = 1 << 20, 148
    counts = wide_cloud_counts(rng, n, sigma)
    start = cta_start_bins(counts, G)
    assert len(start) == G
    assert (counts[start] > 0).all(), "a CTA starts on an empty bin"
    nonempty = np.flatnonzero(counts > 0)
    if len(nonempty) <= G:
        assert set(nonempty) <= set(start.tolist()), "a non-empty bin nobody starts on is only reached by stealing"
    crew = np.bincount(start, minlength=len(counts))
    R = G - min(len(nonempty), G)
    for b in nonempty:
        share = counts[b] / n * R                            # of the CTAs handed out in proportion
        assert abs((crew[b] - (1 if len(nonempty) <= G else 0)) - share) <= 1.0 + 1e-9, (b, crew[b], share)


def test_more_bins_than_ctas():
    rng = np.random.default_rng(7)
    counts = rng.integers(1, 50, 243)
    start = cta_start_bins(counts, 148)
    assert len(start) == 148 and len(set(start.tolist())) == 148     # 148 different bins, the first 148 non-empty ones
    assert (start == np.arange(148)).all()


def test_single_bin():
    counts = np.zeros(243, dtype=int); counts[117] = 1 << 20
    assert (cta_start_bins(counts, 148) == 117).all()
