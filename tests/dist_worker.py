"""Multi-GPU parity worker (run under torchrun, one rank per GPU): the sharded filter must be bit-identical to the
one-GPU filter on the same global particle set -- results, particles, ancestor indices, centroid.

  python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 tests/dist_worker.py
"""
import os
import sys

import numpy as np
import torch
import torch.distributed as td

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from slam_eslam_b200 import GpuPoseEstimator, abi, dist, synthetic  # noqa: E402

KEYS = ("x", "y", "theta", "zpos", "zsigma", "weight", "mprob", "floating", "n_found")


def gather_particles(g, comm):
    p = g.getParticles()
    p["ancestors"] = g.getAncestors()
    out = [None] * comm.world
    td.all_gather_object(out, p)
    return {k: np.concatenate([o[k] for o in out]) for k in p}


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    comm = dist.init(world, rank, local)
    n_local = int(os.environ.get("ESLAM_TEST_N_LOCAL", "8192"))
    n = n_local * world
    for mode in (abi.RESAMPLE_STRATIFIED, abi.RESAMPLE_SYSTEMATIC):
        cfg = dict(min_effective=n + 1, resample_mode=mode)
        g = GpuPoseEstimator(abi.default_config(**cfg), device=local, shard=comm.shard(n_local))
        comm.connect(g)
        contacts, step = synthetic.setup_filter(g, "C2", n=n_local)
        ref = None
        if rank == 0:
            ref = GpuPoseEstimator(abi.default_config(**cfg), device=local)
            synthetic.setup_filter(ref, "C2", n=n)
        for it in range(6):
            if it == 3:
                # skew the weights so that strata of one rank find their ancestors on another rank
                w = np.where(np.arange(n) >= n // 2, 5.0, 0.05) / n
                g.setParticles(weight=w[rank * n_local:(rank + 1) * n_local])
                if ref:
                    ref.setParticles(weight=w)
            r = g.stepFused(step, contacts)
            allp = gather_particles(g, comm)
            if rank == 0:
                rr = ref.stepFused(step, contacts)
                assert r.as_dict() == rr.as_dict(), (it, r.as_dict(), rr.as_dict())
                assert r.resampled == 1
                q = ref.getParticles()
                q["ancestors"] = ref.getAncestors()
                for k in list(KEYS) + ["ancestors"]:
                    assert np.array_equal(allp[k], q[k]), "mode %d step %d: %s differs" % (mode, it, k)
        c = g.getCentroid()
        b = g.getBestParticleIndex()
        if rank == 0:
            assert np.array_equal(c, ref.getCentroid()) and b == ref.getBestParticleIndex()
        comm.barrier()
    # LOCAL_MIGRATE sharding: `world` real shards must reproduce one GPU running the same number of logical shards
    # (cfg.virtual_shards), through local-walk steps and a migration step (a crushed shard)
    cfg = dict(min_effective=n + 1, shard_mode=abi.SHARD_LOCAL_MIGRATE)
    g = GpuPoseEstimator(abi.default_config(**cfg), device=local, shard=comm.shard(n_local))
    comm.connect(g)
    contacts, step = synthetic.setup_filter(g, "C2", n=n_local)
    ref = None
    if rank == 0:
        ref = GpuPoseEstimator(abi.default_config(virtual_shards=world, **cfg), device=local)
        synthetic.setup_filter(ref, "C2", n=n)
    kinds = set()
    for it in range(7):
        if it == 4:
            w = np.where(np.arange(n) < n_local, 1e-4, 1.0) / n      # shard 0 collapses: this step migrates
            g.setParticles(weight=w[rank * n_local:(rank + 1) * n_local])
            if ref:
                ref.setParticles(weight=w)
        r = g.stepFused(step, contacts)
        allp = gather_particles(g, comm)
        if rank == 0:
            rr = ref.stepFused(step, contacts)
            assert r.as_dict() == rr.as_dict(), (it, r.as_dict(), rr.as_dict())
            q = ref.getParticles()
            q["ancestors"] = ref.getAncestors()
            for k in list(KEYS) + ["ancestors"]:
                assert np.array_equal(allp[k], q[k]), "LOCAL_MIGRATE step %d: %s differs" % (it, k)
            kinds.add("local" if np.array_equal(allp["ancestors"] // n_local, np.arange(n) // n_local) else "migrate")
    if rank == 0:
        assert kinds == {"local", "migrate"}, kinds
    comm.barrier()
    # per-particle map patches (cfg.patch_size): the patch travels with its particle, across GPUs too
    P = 8
    cfg = dict(min_effective=n + 1, patch_size=P)
    g = GpuPoseEstimator(abi.default_config(**cfg), device=local, shard=comm.shard(n_local))
    comm.connect(g)
    contacts, step = synthetic.setup_filter(g, "C2", n=n_local)
    ref = None
    if rank == 0:
        ref = GpuPoseEstimator(abi.default_config(**cfg), device=local)
        synthetic.setup_filter(ref, "C2", n=n)
    for it in range(5):
        if it == 2:
            w = np.where(np.arange(n) >= n // 2, 5.0, 0.05) / n       # ancestors (and their patches) cross the shard boundary
            g.setParticles(weight=w[rank * n_local:(rank + 1) * n_local])
            if ref:
                ref.setParticles(weight=w)
        r = g.stepFused(step, contacts)
        allp = gather_particles(g, comm)
        cells, origin = g.getPatches(P)
        outc, outo = [None] * world, [None] * world
        td.all_gather_object(outc, cells); td.all_gather_object(outo, origin)
        if rank == 0:
            rr = ref.stepFused(step, contacts)
            assert r.as_dict() == rr.as_dict(), (it, r.as_dict(), rr.as_dict())
            q = ref.getParticles(); q["ancestors"] = ref.getAncestors()
            for k in list(KEYS) + ["ancestors"]:
                assert np.array_equal(allp[k], q[k]), "patches step %d: %s differs" % (it, k)
            rc, ro = ref.getPatches(P)
            assert np.array_equal(np.concatenate(outo), ro), "patch origins differ at step %d" % it
            assert np.array_equal(np.concatenate(outc), rc, equal_nan=True), "patch cells differ at step %d" % it
            if it == 4:
                assert (~np.isnan(rc[..., 0])).sum() > n // 2
    comm.barrier()
    # a legal weight vector whose exact running sum overflows the boundary pool: the sequential fallback runs rank after rank
    # (each continues from the sum the previous rank ended with) and N ranks still equal one GPU and numpy's sequential sum
    g = GpuPoseEstimator(abi.default_config(), device=local, shard=comm.shard(n_local))
    comm.connect(g)
    rng = np.random.default_rng(3)
    w = np.full(n, 1e-17); w[0] = 1.0
    w[1::2] = 2.0 ** -53 * (1 + rng.random(n // 2))
    z = np.zeros(n_local)
    g.setParticles(x=np.arange(rank * n_local, (rank + 1) * n_local, dtype=float), y=z, theta=z, zpos=z, zsigma=z + 1,
                   weight=w[rank * n_local:(rank + 1) * n_local])
    g.resample()
    anc, cdf, fb = [None] * world, [None] * world, [None] * world
    td.all_gather_object(anc, g.getAncestors()); td.all_gather_object(cdf, g.getCdf()); td.all_gather_object(fb, g.debugCounters())
    if rank == 0:
        ref = GpuPoseEstimator(abi.default_config(), device=local)
        zz = np.zeros(n)
        ref.setParticles(x=np.arange(n, dtype=float), y=zz, theta=zz, zpos=zz, zsigma=zz + 1, weight=w)
        ref.resample()
        assert np.array_equal(np.concatenate(cdf), np.cumsum(w)), "sharded sequential fallback: running sum"
        assert np.array_equal(np.concatenate(anc), ref.getAncestors()), "sharded sequential fallback: ancestors"
        assert all(f["fallbacks"] == 1 and f["error_flags"] == 0 for f in fb), fb
    comm.barrier()
    if rank == 0:
        print("dist parity ok: %d ranks x %d particles bit-identical to one GPU (exact-global, systematic, LOCAL_MIGRATE, patches, sequential fallback)" % (world, n_local))
    td.destroy_process_group()


if __name__ == "__main__":
    main()
