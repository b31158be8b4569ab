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
    if rank == 0:
        print("dist parity ok: %d ranks x %d particles bit-identical to one GPU (exact-global, systematic, LOCAL_MIGRATE)" % (world, n_local))
    td.destroy_process_group()


if __name__ == "__main__":
    main()
