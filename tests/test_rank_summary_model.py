"""What the ranks of a sharded filter exchange per step (eslam_kernels.cuh "peer exchange"), restated in numpy /
Python and checked on the CPU:

  * sums: a shard holds 2^k tiles, so the index-aligned pairwise tree over all G T tiles factors into G per-rank subtrees
    whose roots are combined by the same tree rule -- every rank reduces its own tiles and exchanges ONE value, and the
    result has the bits of the one-GPU tree (block_tree_reduce / tree_small);
  * exact running sum: a rank summarises its part of the walk as (head function, [boundary weight, function of the run
    after it]...); walking the summaries of the ranks in front and one's own reproduces the sequential fp64 running sum
    (src/ParticleFilter.hpp:95,102) at every rank start and every boundary element (k_chain / chain_walk);
and, with gloo and world_size 2, the same exchange between two real processes.
"""
import os
import subprocess
import sys

import numpy as np
import pytest

from test_exact_sum_model import BIN_NONE, apply, approximate_prefix, classify, compose, vectors

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def tree(v):
    """block_tree_reduce / tree_small: pairwise over the index-aligned binary tree, a missing right operand passes through."""
    v = [float(x) for x in v]
    while len(v) > 1:
        v = [v[2 * j] + v[2 * j + 1] if 2 * j + 1 < len(v) else v[2 * j] for j in range((len(v) + 1) // 2)]
    return v[0]


@pytest.mark.parametrize("G", [2, 3, 4, 5, 6, 7, 8])
@pytest.mark.parametrize("T", [1, 2, 8, 64, 1024])
def test_tree_over_all_tiles_factors_into_rank_subtrees(G, T):
    rng = np.random.default_rng(100 * G + T)
    for scale in (1.0, 1e-12, 1e9):
        tiles = rng.random(G * T) * scale * 10.0 ** rng.integers(-6, 6, G * T)
        one_gpu = tree(tiles)
        roots = [tree(tiles[r * T:(r + 1) * T]) for r in range(G)]
        assert tree(roots) == one_gpu


def test_factorisation_needs_power_of_two_shards():
    """With 3 tiles per rank the first tree level pairs tile 2 of rank 0 with tile 0 of rank 1: the library refuses such shards."""
    rng = np.random.default_rng(5)
    differs = 0
    for _ in range(50):
        tiles = rng.random(6)
        differs += tree([tree(tiles[:3]), tree(tiles[3:])]) != tree(tiles)
    assert differs > 0


def rank_summary(w, A_prev_of_first, A, eps):
    """k_norm_scan + k_chain for one rank: classify every element from the approximate prefix, return
    (head function, [(boundary weight, function of the run after it)], bad)."""
    head, entries, bad = (BIN_NONE, 0, 0), [], False
    for i in range(len(w)):
        kind, b, d0, d1 = classify(float(w[i]), float(A[i - 1]) if i else float(A_prev_of_first), float(A[i]), eps)
        if kind == 2:
            entries.append([float(w[i]), (BIN_NONE, 0, 0)])
        elif kind == 1:
            if entries:
                entries[-1][1], bd = compose(entries[-1][1], (b, d0, d1))
            else:
                head, bd = compose(head, (b, d0, d1))
            bad |= bd
    return head, entries, bad


def chain_walk(S, head, entries):
    """chain_walk of k_chain; also returns the sum at every boundary element."""
    S, bad = apply(S, head)
    at = []
    for wk, f in entries:
        S = S + wk
        at.append(S)
        S, b2 = apply(S, f)
        bad |= b2
    return S, at, bad


@pytest.mark.parametrize("G", [2, 4, 8])
def test_rank_summaries_reproduce_the_sequential_running_sum(G):
    n_local = 4096
    n = G * n_local
    rng = np.random.default_rng(G)
    eps = max(8 * n * 2.0 ** -53, 2.0 ** -40)
    for name, w in vectors(n, rng):
        A = approximate_prefix(w)
        seq = np.cumsum(w)
        summaries = []
        for r in range(G):
            lo, hi = r * n_local, (r + 1) * n_local
            summaries.append(rank_summary(w[lo:hi], A[lo - 1] if lo else 0.0, A[lo:hi], eps))
        assert not any(s[2] for s in summaries), name
        S = 0.0
        n_bnd = 0
        for r in range(G):                                   # what rank G-1 does; every rank stops after its own summary
            head, entries, _ = summaries[r]
            S, at, bad = chain_walk(S, head, entries)
            assert not bad, (name, r)
            assert S == seq[(r + 1) * n_local - 1], "%s: running sum at the end of rank %d" % (name, r)
            n_bnd += len(entries)
        assert n_bnd < 400, (name, n_bnd)                    # O(1) per rank is what travels


WORKER = r'''
import os, sys
sys.path.insert(0, %r); sys.path.insert(0, os.path.join(%r, "tests"))
import numpy as np, torch, torch.distributed as td
from slam_eslam_b200 import dist
from test_rank_summary_model import tree, rank_summary, chain_walk
from test_exact_sum_model import approximate_prefix
rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
comm = dist.init(world, rank, 0, backend="gloo")
uid = comm.broadcast_bytes(bytes(range(128)) if rank == 0 else b"", 128)
assert uid == bytes(range(128)), uid
assert comm.max_float(1.5 + rank) == 1.5 + world - 1
s = comm.shard(4096)
assert (s.rank, s.world, s.global_count) == (rank, world, 4096 * world)
# the filter's exchanges between two real processes: each rank owns 4 tiles of 1024 weights
T, n_local = 4, 4096
rng = np.random.default_rng(7)
w = rng.random(world * n_local) ** 3; w /= w.sum()
mine = w[rank * n_local:(rank + 1) * n_local]
root = tree([tree(mine[t * 1024:(t + 1) * 1024]) for t in range(T)])          # the rank's subtree root (last CTA of k_loop_b)
roots = [None] * world
td.all_gather_object(roots, root)                                              # exchange X_B: one double per rank
tiles = [tree(w[t * 1024:(t + 1) * 1024]) for t in range(world * T)]
assert tree(roots) == tree(tiles), "the combined tree differs from the one-GPU tree"
A = approximate_prefix(w)
eps = max(8 * len(w) * 2.0 ** -53, 2.0 ** -40)
lo = rank * n_local
summ = rank_summary(mine, A[lo - 1] if lo else 0.0, A[lo:lo + n_local], eps)
allsum = [None] * world
td.all_gather_object(allsum, summ)                                             # exchange X_C: the rank's chain summary
S = 0.0
for r in range(rank + 1):
    S, at, bad = chain_walk(S, allsum[r][0], allsum[r][1])
    assert not bad
assert S == np.cumsum(w)[lo + n_local - 1], "rank %%d: running sum at its end" %% rank
comm.barrier()
print("rank %%d ok" %% rank)
'''


def test_exchanges_between_two_processes_gloo_world2(tmp_path):
    script = tmp_path / "worker.py"
    script.write_text(WORKER % (ROOT, ROOT))
    env = dict(os.environ, MASTER_ADDR="127.0.0.1", MASTER_PORT="29611")
    out = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2",
                          "--master-addr", "127.0.0.1", "--master-port", "29611", str(script)],
                         env=env, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, timeout=240)
    assert out.returncode == 0, out.stdout[-3000:]
    assert "rank 0 ok" in out.stdout and "rank 1 ok" in out.stdout
