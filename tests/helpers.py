"""Shared helpers for the parity tests."""
import json
import math
import os

import numpy as np

from slam_eslam_b200 import abi

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def load_kat():
    with open(os.path.join(GOLDEN, "contact_model_kat.json")) as f:
        return json.load(f)


def kat_contacts(case):
    contact = [float(c) for c in case["contact"]]  # "nan" -> nan (base::unknown<float>())
    return abi.make_contacts(case["points"], contact, case["group"])


def quadrant_grid(z, stddev, res=None):
    """The 4-quadrant FakeMLSAccess as a 2x2 dense grid (cells 1000 m wide, origin at the centre):
    cell (ix, iy) = (x >= 0, y >= 0) holds z/stddev[2*iy + ix]; res[i] == False -> empty cell."""
    cells = np.zeros((2, 2, 2), dtype=np.float32)
    for iy in range(2):
        for ix in range(2):
            i = 2 * iy + ix
            cells[iy, ix, 0] = np.nan if (res is not None and not res[i]) else z[i]
            cells[iy, ix, 1] = stddev[i]
    return cells, (1000.0, 1000.0), (-1000.0, -1000.0)


def kat_config(case, **kw):
    return abi.default_config(contact_point_radius=case.get("radius", 0.01),
                              min_contacts=case.get("min_contacts", 3), **kw)


def run_trajectory(f, contacts, step, n_steps, collect=("weight",)):
    """project + update for n_steps; returns the per-step UpdateResult dicts and final particles."""
    results = []
    for _ in range(n_steps):
        f.project(step)
        results.append(f.update(contacts).as_dict())
    return results, f.getParticles()


def assert_particles_equal(a, b, exact=True, rtol=0.0, atol=0.0, keys=None):
    keys = keys or ("x", "y", "theta", "zpos", "zsigma", "weight", "mprob", "floating", "n_found")
    for k in keys:
        if exact or a[k].dtype == np.uint8:
            bad = np.flatnonzero(~((a[k] == b[k]) | (np.isnan(a[k].astype(float)) & np.isnan(b[k].astype(float)))))
            assert bad.size == 0, "%s differs at %d particles, first %d: %r vs %r" % (
                k, bad.size, bad[0], a[k][bad[0]], b[k][bad[0]])
        else:
            np.testing.assert_allclose(a[k], b[k], rtol=rtol, atol=atol, err_msg=k)
