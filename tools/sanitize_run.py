#!/usr/bin/env python
"""Small hot-path run for compute-sanitizer (memcheck / racecheck / synccheck / initcheck): C1 and C2 with forced
resampling, the generic kernel with per-particle patches, centroid and best particle.  Checked against the oracle so that
a sanitizer-clean run is also a correct one.

  compute-sanitizer --tool racecheck python tools/sanitize_run.py
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import binding  # noqa: E402  (the checker)
from slam_eslam_b200 import GpuPoseEstimator, abi, synthetic  # noqa: E402


def run(workload, n, steps, **cfg):
    c = dict(min_effective=n + 1, **cfg)
    g = GpuPoseEstimator(abi.default_config(**c), device=0)
    o = binding.OraclePoseEstimator(abi.default_config(**c), policy=binding.SHARED)
    contacts, step = synthetic.setup_filter(g, workload, n=n)
    synthetic.setup_filter(o, workload, n=n)
    for i in range(steps):
        ra = g.stepFused(step, contacts)
        o.project(step)
        rb = o.update(contacts)
        assert ra.as_dict() == rb.as_dict(), (workload, i)
        assert np.array_equal(g.getAncestors(), o.getAncestors()), (workload, i)
    assert np.array_equal(g.getCentroid(), o.getCentroid())
    assert g.getBestParticleIndex() == o.getBestParticleIndex()
    g.close()
    print("ok", workload, n, cfg)


if __name__ == "__main__":
    run("C1", 250, 3)
    run("C2", 10000, 3)
    run("C2", 3000, 2, resample_mode=abi.RESAMPLE_SYSTEMATIC)
    run("C1", 2500, 2, patch_size=8)
    run("C2", 8192, 2, shard_mode=abi.SHARD_LOCAL_MIGRATE, virtual_shards=4)
