"""N > 1 path on the CPU: (1) the host-side planning step of sharded resampling (product code, no device needed)
against the oracle's ancestors; (2) the rendezvous plumbing over gloo with world_size 2."""
import ctypes as C
import os
import subprocess
import sys

import numpy as np
import pytest

from slam_eslam_b200 import abi, build

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
TILE = 1024


def plan(lib, tile_start, n_local, world):
    b = (C.c_uint64 * world)()
    e = (C.c_uint64 * world)()
    ts = np.ascontiguousarray(tile_start, dtype=np.float64)
    st = lib.eslam_gpu_plan_windows(ts.ctypes.data_as(C.POINTER(C.c_double)), len(ts), n_local, world, b, e)
    assert st == abi.OK
    return list(b), list(e)


@pytest.mark.parametrize("world", [2, 4, 8])
@pytest.mark.parametrize("kind", ["uniform", "skewed", "one heavy", "zeros first"])
def test_window_plan_contains_every_ancestor(oracle, world, kind):
    lib = abi.load_library(build.build())
    n_local = 2 * TILE
    n = n_local * world
    rng = np.random.default_rng(world)
    if kind == "uniform":
        w = np.full(n, 1.0 / n)
    elif kind == "skewed":
        w = np.where(np.arange(n) >= n // 2, 5.0, 0.05) * rng.random(n)
        w /= w.sum()
    elif kind == "one heavy":
        w = rng.random(n) * 1e-9
        w[n // 3] = 1.0
    else:
        w = rng.random(n)
        w[: n // 2 + 17] = 0.0
        w /= w.sum()
    f = oracle.OraclePoseEstimator(abi.default_config(), policy=oracle.LITERAL)
    z = np.zeros(n)
    f.setParticles(x=z, y=z, theta=z, zpos=z, zsigma=z + 1, weight=w)
    f.resample()
    anc = f.getAncestors().astype(np.int64)
    cdf = np.cumsum(w)                                             # sequential, like the reference
    tile_start = np.concatenate(([0.0], cdf[TILE - 1:-1:TILE]))    # exact running sum before each tile
    begin, end = plan(lib, tile_start, n_local, world)
    for r in range(world):
        a = anc[r * n_local:(r + 1) * n_local]
        assert begin[r] % TILE == 0 and begin[r] <= a.min() and a.max() < end[r] <= n, (r, begin[r], end[r], a.min(), a.max())
    if kind == "uniform":                                          # near-uniform weights: own range +- one tile
        for r in range(world):
            assert begin[r] >= max(0, r * n_local - TILE) and end[r] <= min(n, (r + 1) * n_local + TILE)


WORKER = r'''
import os, sys
sys.path.insert(0, %r)
import torch
from slam_eslam_b200 import dist
rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
comm = dist.init(world, rank, 0, backend="gloo")
uid = comm.broadcast_bytes(bytes(range(128)) if rank == 0 else b"", 128)
assert uid == bytes(range(128)), uid
assert comm.max_float(1.5 + rank) == 1.5 + world - 1
s = comm.shard(4096)
assert (s.rank, s.world, s.global_count) == (rank, world, 4096 * world)
comm.barrier()
print("rank %%d ok" %% rank)
'''


def test_rendezvous_plumbing_gloo_world2(tmp_path):
    script = tmp_path / "worker.py"
    script.write_text(WORKER % ROOT)
    env = dict(os.environ, MASTER_ADDR="127.0.0.1", MASTER_PORT="29611")
    out = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2",
                          "--master-addr", "127.0.0.1", "--master-port", "29611", str(script)],
                         env=env, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, timeout=240)
    assert out.returncode == 0, out.stdout[-3000:]
    assert "rank 0 ok" in out.stdout and "rank 1 ok" in out.stdout
