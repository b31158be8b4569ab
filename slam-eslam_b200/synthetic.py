"""Synthetic terrain, contact points and odometry streams for the configurations of BASELINE.json
(SURVEY.md §8d).  Inputs only -- nothing here is on the measured path."""
import functools
import math

import numpy as np

from . import abi


@functools.lru_cache(maxsize=4)
def make_terrain(width, height, scale=0.05, seed=1234, empty_fraction=0.01):
    """Dense elevation grid: mean = 0.25 sin(2 pi x/5) cos(2 pi y/7) + 0.05 n, stdev = 0.01 + 0.02 u,
    `empty_fraction` of the cells empty (NaN mean).  Returns (cells[h,w,2] float32, scale(2), offset(2))."""
    rng = np.random.Generator(np.random.Philox(seed))
    ox, oy = -width * scale / 2.0, -height * scale / 2.0
    xs = ox + (np.arange(width) + 0.5) * scale
    ys = oy + (np.arange(height) + 0.5) * scale
    X, Y = np.meshgrid(xs, ys)
    mean = 0.25 * np.sin(2 * np.pi * X / 5.0) * np.cos(2 * np.pi * Y / 7.0)
    mean = mean + 0.05 * rng.standard_normal(mean.shape)
    stdev = 0.01 + 0.02 * rng.random(mean.shape)
    empty = rng.random(mean.shape) < empty_fraction
    cells = np.empty((height, width, 2), dtype=np.float32)
    cells[..., 0] = mean.astype(np.float32)
    cells[..., 1] = stdev.astype(np.float32)
    cells[..., 0][empty] = np.nan
    return cells, (scale, scale), (ox, oy)


def terrain_height(x, y):
    return 0.25 * math.sin(2 * math.pi * x / 5.0) * math.cos(2 * math.pi * y / 7.0)


def make_contact_points(n_points, wheel_phase=0.0):
    """n_points in 4 contiguous groups (groupId 0..3) on 4 wheels at (+-0.25, +-0.25) m, each group's points on
    a radius-0.18 m circle in the x-z plane (Asguard's star wheels); contact = 0.5 except every 8th point 0.1
    (below the 0.2 gate, src/ContactModel.cpp:136) and one NaN (unknown -> passes the gate)."""
    assert n_points % 4 == 0
    per = n_points // 4
    pos, contact, gid = [], [], []
    centres = [(0.25, 0.25), (0.25, -0.25), (-0.25, 0.25), (-0.25, -0.25)]
    k = 0
    for g, (cx, cy) in enumerate(centres):
        for j in range(per):
            a = wheel_phase + 2 * math.pi * j / per - math.pi / 2 + 0.3 * g
            pos.append((cx + 0.18 * math.cos(a), cy, 0.18 * math.sin(a)))
            contact.append(0.1 if (k % 8) == 7 else 0.5)
            gid.append(g)
            k += 1
    contact[1] = float("nan")
    return abi.make_contacts(pos, contact, gid)


def odometry_step(yaw=0.0):
    """Constant odometry stream of SURVEY.md §8d: mean delta (0.02 m, 0, 0.005 rad), sigma (0.002, 0.002, 0.001)."""
    return abi.make_step(yaw=yaw, z_delta=0.0, z_var=2e-6, delta_mean=(0.02, 0.0, 0.005),
                         delta_sigma=(0.002, 0.002, 0.001))


WORKLOADS = {
    # name: particles, contact points, grid, sigma_xy of the initial cloud
    "C1": dict(n=250, c=20, grid=200, sigma_xy=0.1),
    "C2": dict(n=10000, c=16, grid=512, sigma_xy=0.1),
    "C3": dict(n=1048576, c=64, grid=2048, sigma_xy=2.0),
    "C4": dict(n=100000, c=64, grid=4096, sigma_xy=2.0),
    "C5": dict(n=16777216, c=64, grid=4096, sigma_xy=2.0),
    "C3x8": dict(n=8388608, c=64, grid=2048, sigma_xy=2.0),   # 8192 tiles on one GPU: what the single-CTA kernels see at 8 ranks
}


def init_cloud(f, workload, n=None):
    """The initial cloud of a workload (again): same seed, same particles."""
    w = WORKLOADS[workload]
    n = n or w["n"]
    f.init(n, (0.0, 0.0, 0.0), (w["sigma_xy"], w["sigma_xy"], 0.1), terrain_height(0, 0) + 0.19, 1.001)


def setup_filter(f, workload, n=None, force_resample=True):
    """Attach terrain and draw the initial cloud on any filter object exposing the PoseEstimator mirror API."""
    w = WORKLOADS[workload]
    n = n or w["n"]
    cells, scale, offset = make_terrain(w["grid"], w["grid"])
    f.setEnvironment(cells, scale, offset)
    init_cloud(f, workload, n)
    return make_contact_points(w["c"]), odometry_step()
