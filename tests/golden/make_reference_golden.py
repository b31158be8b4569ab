#!/usr/bin/env python
"""Generates tests/golden/reference_trajectories.npz from the REFERENCE'S OWN compiled code (oracle/_ref/libeslam_ref.so:
/root/reference/src/{ContactModel,PoseEstimator}.cpp + headers, unmodified, against the stand-in dependency headers of
oracle/ref_shim; see oracle/ref/ref_driver.cpp).  Run in the build container (needs /root/reference):

    python tests/golden/make_reference_golden.py

Cases (synthetic inputs of slam_eslam_b200.synthetic, Philox noise streams, seed 42):
  c1_default : C1 (250 particles x 20 contact points, 200^2 grid), reference defaults (minEffective 50), 30 steps
  c1_forced  : the same with minEffective = 251 (stratified resampling every step), 30 steps
  c2_forced  : C2 (10 000 particles x 16 contact points, 512^2 grid), resampling every step, 6 steps
Stored per case: every particle field after the steps listed in `<case>/steps`, the ancestors of every resampling step,
the ESS series (NaN where the reference resampled: update() does not return it), centroid and best index at the end.
For the 10 000-particle case the arrays are stored as SHA-256 digests of their bytes (bit-exactness is what is checked).
"""
import hashlib
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import ref_binding  # noqa: E402
from slam_eslam_b200 import abi, synthetic  # noqa: E402

CASES = {
    "c1_default": dict(workload="C1", min_effective=50, steps=30, keep=(0, 1, 5, 19, 20, 21, 29)),
    "c1_forced": dict(workload="C1", min_effective=251, steps=30, keep=(0, 1, 10, 29)),
    "c2_forced": dict(workload="C2", min_effective=10001, steps=6, keep=(0, 5), digest=True),
}
FIELDS = ("x", "y", "theta", "zpos", "zsigma", "weight", "mprob", "floating", "n_found")


def digest(a):
    return np.array(hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest())


def run_case(make_filter, spec):
    """Shared by the generator and by tests/test_reference_parity.py (which runs the oracle through it)."""
    store = digest if spec.get("digest") else (lambda a: a.copy())
    f = make_filter(abi.default_config(min_effective=spec["min_effective"]))
    contacts, step = synthetic.setup_filter(f, spec["workload"])
    out = {"steps": np.array(spec["keep"]), "ess": [], "resampled": [], "data_particles": []}
    anc = []
    for i in range(spec["steps"]):
        f.project(step)
        r = f.update(contacts)
        out["ess"].append(r.ess if not r.resampled else float("nan"))
        out["resampled"].append(r.resampled)
        if r.resampled:
            anc.append(f.getAncestors().copy())
        if i in spec["keep"]:
            p = f.getParticles()
            for k in FIELDS:
                out["%s_%d" % (k, i)] = store(p[k])
    out["ancestors"] = store(np.array(anc, dtype=np.uint32).reshape(len(anc), -1))
    out["ess"] = np.array(out["ess"]); out["resampled"] = np.array(out["resampled"], dtype=np.int32)
    out["centroid"] = f.getCentroid()
    out["best_index"] = np.array([f.getBestParticleIndex()], dtype=np.uint64)
    out.pop("data_particles")
    return out


def main():
    ref_binding.build(force=True)
    blob = {}
    for name, spec in CASES.items():
        res = run_case(lambda cfg: ref_binding.RefPoseEstimator(cfg, debug=True), spec)
        for k, v in res.items():
            blob["%s/%s" % (name, k)] = v
        print(name, "resampling steps:", int(res["resampled"].sum()), "last ess:", res["ess"][-1])
    path = os.path.join(ROOT, "tests", "golden", "reference_trajectories.npz")
    np.savez_compressed(path, **blob)
    print("wrote", path, os.path.getsize(path), "bytes")


if __name__ == "__main__":
    main()
