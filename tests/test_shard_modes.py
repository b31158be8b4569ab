"""Sharded resampling modes of the filter (include/eslam_gpu.h ESLAM_SHARD_*; SURVEY.md §8e, BASELINE config 5:
"NCCL weight allreduce and ESS-triggered migration"), on the oracle: it holds all particles and treats them as G logical
shards, which is also how one GPU runs the mode (cfg.virtual_shards) and what G real shards must reproduce.

EXACT_GLOBAL is the reference's own walk (pinned in test_reference_parity.py).  LOCAL_MIGRATE has no reference code behind
it: these tests state its definition -- local walks over each shard's own running sum, a global walk (the migration) in any
step where a shard's weight share or own ESS has collapsed."""
import numpy as np
import pytest

from slam_eslam_b200 import abi, synthetic

N, G = 8192, 4
NS = N // G


def make(oracle, policy, **cfg):
    o = oracle.OraclePoseEstimator(abi.default_config(min_effective=N + 1, **cfg), policy=policy)
    contacts, step = synthetic.setup_filter(o, "C2", n=N)
    return o, contacts, step


def run(oracle, policy, steps=6, **cfg):
    o, contacts, step = make(oracle, policy, **cfg)
    anc = []
    for _ in range(steps):
        o.project(step)
        assert o.update(contacts).resampled == 1
        anc.append(o.getAncestors().copy())
    return anc, o.getParticles()


@pytest.mark.parametrize("policy", [0, 1])
def test_local_walks_stay_inside_their_shard(oracle, policy):
    anc, _ = run(oracle, policy, shard_mode=abi.SHARD_LOCAL_MIGRATE, virtual_shards=G)
    exact, _ = run(oracle, policy)
    own = np.arange(N) // NS
    for a in anc:
        assert np.array_equal(a // NS, own)                       # no particle crossed a shard boundary
        for g in range(G):
            assert np.all(np.diff(a[g * NS:(g + 1) * NS].astype(np.int64)) >= 0)    # each local walk is monotone
    assert any(not np.array_equal(a, e) for a, e in zip(anc, exact))               # and it is not the global walk


@pytest.mark.parametrize("policy", [0, 1])
def test_one_shard_and_forced_migration_equal_the_exact_walk(oracle, policy):
    exact, pe = run(oracle, policy)
    for cfg in (dict(shard_mode=abi.SHARD_LOCAL_MIGRATE, virtual_shards=1),
                dict(shard_mode=abi.SHARD_LOCAL_MIGRATE, virtual_shards=G, migrate_min_share=1e9),     # every step migrates
                dict(shard_mode=abi.SHARD_EXACT_GLOBAL, virtual_shards=G)):                           # mode off
        anc, p = run(oracle, policy, **cfg)
        assert all(np.array_equal(a, e) for a, e in zip(anc, exact)), cfg
        assert all(np.array_equal(p[k], pe[k]) for k in p)


@pytest.mark.parametrize("policy", [0, 1])
def test_migration_triggers_on_share_and_on_ess_collapse(oracle, policy):
    zeros = np.zeros(N)
    base = dict(x=np.arange(N, dtype=float), y=zeros, theta=zeros, zpos=zeros, zsigma=zeros + 1)

    def resample_with(weights, **cfg):
        o = oracle.OraclePoseEstimator(abi.default_config(**cfg), policy=policy)
        o.setParticles(weight=weights, **base)
        o.normalizeWeights()
        o.resample()
        return o.getAncestors()
    local = dict(shard_mode=abi.SHARD_LOCAL_MIGRATE, virtual_shards=G)
    rng = np.random.default_rng(1)
    own = np.arange(N) // NS
    # balanced weights: local walks
    w = rng.random(N) + 0.5
    assert np.array_equal(resample_with(w, **local) // NS, own)
    # shard 2 holds 1% of a fair share: migration = the exact-global walk, its slots are refilled from the other shards
    w2 = w.copy(); w2[2 * NS:3 * NS] *= 0.01
    a = resample_with(w2, **local)
    assert np.array_equal(a, resample_with(w2)) and not np.array_equal(a // NS, own)
    # shard 1 has a fair share but all of it on 3 particles: its own ESS collapses
    w3 = w.copy(); w3[NS:2 * NS] = 1e-9; w3[NS + 5] = w3[NS + 700] = w3[NS + 1500] = w[:NS].sum() / 3
    a = resample_with(w3, **local)
    assert np.array_equal(a, resample_with(w3))
    # ... unless the thresholds allow it: then shard 1 resamples its three heavy particles itself
    a = resample_with(w3, migrate_min_ess=0.0, **local)
    assert np.array_equal(a // NS, own) and set(a[NS:2 * NS]) <= {NS + 5, NS + 700, NS + 1500, 2 * NS - 1}
    # all weights zero: uniform reset (Q8), nothing to migrate
    assert np.array_equal(resample_with(zeros, **local) // NS, own)


def test_local_strata_use_the_shard_total(oracle):
    """Stratum k of shard g draws (k + u)/n_s * W_g: every particle of a shard with uniform weights is its own ancestor
    whatever share W_g of the total the shard holds."""
    w = np.repeat([1.0, 2.0, 0.5, 1.5], NS)
    zeros = np.zeros(N)
    o = oracle.OraclePoseEstimator(abi.default_config(shard_mode=abi.SHARD_LOCAL_MIGRATE, virtual_shards=G), policy=1)
    o.setParticles(x=zeros, y=zeros, theta=zeros, zpos=zeros, zsigma=zeros + 1, weight=w)
    o.normalizeWeights(); o.resample()
    a = o.getAncestors().astype(np.int64)
    assert np.abs(a - np.arange(N)).max() <= 1
