#!/usr/bin/env python
"""Benchmark of the eSLAM particle-filter update path (BASELINE.json metric).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference] [--workload C3]

A "step" is one pass of the hot path over the whole particle set: project -> contact weight update ->
normalise/ESS -> stratified resample (forced every step) -> gather.  metric = particle-updates/s =
N particles x C contact points x steps / time.  `value` is measured with inputs resident in HBM (CUDA events on the
library's stream, max over ranks); `e2e` goes through the reference-facing C-ABI call with host buffers each step.
"""
import argparse
import ctypes
import json
import os
import subprocess
import sys
import threading
import time

# stdout carries exactly one JSON line: file descriptor 1 is pointed at stderr for the whole run (NCCL prints its version banner to
# stdout whatever NCCL_DEBUG_FILE says), and the line is written to the saved descriptor at the end
_REAL_STDOUT = os.dup(1)
os.dup2(2, 1)


def emit(obj):
    sys.stdout.flush()
    os.write(_REAL_STDOUT, (json.dumps(obj) + "\n").encode())

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)



def ncu_traffic(workload, kernel):
    """dram__bytes_read.sum + dram__bytes_write.sum of one launch of the dominant kernel: not measurable inside a timed run
    (counters need ncu's replay), so it is read from the summary tools/ncu_traffic.py wrote from the committed
    `ncu --set full` capture of this same command (profiles/<round>/traffic.json); null when there is none."""
    for rnd in ("r02", "r01"):
        p = os.path.join(ROOT, "profiles", rnd, "traffic.json")
        if os.path.exists(p):
            try:
                t = json.load(open(p)).get("%s:%s" % (workload, kernel))
                if t:
                    return {"bytes": float(t["bytes"]), "source": "profiles/%s/%s" % (rnd, t.get("source", "traffic.json"))}
            except Exception:
                pass
    return None

KIND_TEXT = {"reference": "the reference's own PoseEstimator / ContactModel sources compiled against stand-in dependency headers "
                          "(dense-grid map, Philox noise)",
             "port": "LITERAL oracle port (oracle/_ref not built)"}
METRIC = "particle-updates/sec"
UNIT = "particle-updates/s"


def algorithmic_bytes(n, c, lean, patch=0):
    """B_step(N, C) = N (8 C + 234); with per-particle P x P map patches (config C4) + 4 * 16 (read-modify-write
    of at most 4 patch cells in the update) + 2 * P^2 * 8 (the patch read and written when the particle is resampled).
    The dominant kernel's share: k_weight_lean reads the C map cells (8 C), the particle's x, y, theta, z, zSigma,
    weight (48) and writes z, zSigma, weight, mprob, floating, n_found (34): N (8 C + 82); the generic fused project+weight
    kernel moves §8(d)'s N (8 C + 97), + 4 * 16 with patches."""
    per_patch = (4 * 16 + 2 * patch * patch * 8) if patch else 0
    return n * (8 * c + 234 + per_patch), n * (8 * c + (82 if lean else 97) + (4 * 16 if patch else 0))


def measured_peak():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured"
        except Exception:
            pass
    return 6650.0, "fallback"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled during the timed region (B200_PROFILING.md recipe)."""
    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.rows, self.proc = [], None
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(index), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "20"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([f.strip() for f in line.split(",")])

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], None, set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            try:
                sm.append(float(r[0])); mx = float(r[1])
            except Exception:
                continue
            for name, v in zip(names, r[2:6]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": mx, "reasons": sorted(reasons),
                "samples": len(sm)}


def cpu_kind():
    """"reference": oracle/_ref/libeslam_ref.so, the reference's own src/{ContactModel,PoseEstimator}.cpp compiled unmodified
    against stand-in dependency headers (oracle/ref/, oracle/ref_shim/; prebuilt in the build container, it travels with the
    snapshot).  "port": the LITERAL oracle, when that library is not there."""
    from oracle import ref_binding
    return "reference" if ref_binding.available() else "port"


def _oracle_shard(workload, n, steps, warmup, out, k):
    """One independent CPU filter of n particles (the C call releases the GIL, so shards run on separate cores)."""
    from oracle import binding, ref_binding
    from slam_eslam_b200 import abi, synthetic
    cfg = abi.default_config(min_effective=n + 1)
    if cpu_kind() == "reference":
        f = ref_binding.RefPoseEstimator(cfg, debug=False)
    else:
        f = binding.OraclePoseEstimator(cfg, policy=binding.LITERAL)
    contacts, step = synthetic.setup_filter(f, workload, n=n)
    for _ in range(warmup):
        f.project(step); f.update(contacts)
    out[k] = (f, contacts, step)


def time_oracle(workload, n_total, steps, warmup, threads):
    """The reference's CPU filter (its own compiled sources when oracle/_ref is there, else the LITERAL oracle port: the two
    are bit-identical, tests/test_reference_parity.py) on `threads` host threads.  The reference itself is single-threaded
    (its only OpenMP loop is racy, src/PoseEstimator.cpp:272-327), so threads > 1 runs that many independent filters of
    n_total/threads particles side by side: an upper bound for what a particle-parallel CPU build of it could do.  The map
    behind its lookups is the dense stand-in grid, cheaper than envire's MLS map.  Returns particle-updates/s."""
    from slam_eslam_b200 import synthetic
    w = synthetic.WORKLOADS[workload]
    n = max(1, n_total // threads)
    synthetic.make_terrain(w["grid"], w["grid"])       # cached: generated once, shared by the shards
    shards = [None] * threads
    ts = [threading.Thread(target=_oracle_shard, args=(workload, n, steps, warmup, shards, k)) for k in range(threads)]
    [t.start() for t in ts]; [t.join() for t in ts]

    def work(k):
        f, contacts, step = shards[k]
        for _ in range(steps):
            f.project(step); f.update(contacts)
    ts = [threading.Thread(target=work, args=(k,)) for k in range(threads)]
    t0 = time.perf_counter()
    [t.start() for t in ts]; [t.join() for t in ts]
    dt = time.perf_counter() - t0
    return n * threads * w["c"] * steps / dt, dt, n


def run_reference(args, rank, world):
    """--impl reference: the reference's CPU filter on this box's host cores -- eslam::PoseEstimator::project + update from
    the reference's own sources (oracle/_ref/libeslam_ref.so, see cpu_kind), on all host threads."""
    if rank != 0:
        return
    from slam_eslam_b200 import synthetic
    w = synthetic.WORKLOADS[args.workload]
    threads = max(1, min(os.cpu_count() or 1, 64))     # every filter holds its own copy of the map
    n_sample = min(w["n"], args.cpu_particles * threads)
    value, dt, n = time_oracle(args.workload, n_sample, args.steps, args.warmup, threads)
    kind = cpu_kind()
    sample = ("%d independent CPU filters x %d particles x %d contact points per step (%d of the workload's %d particles; same "
              "terrain, contacts, odometry; forced resampling); %s" % (threads, n, w["c"], n * threads, w["n"], KIND_TEXT[kind]))
    emit(({
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * dt / args.steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": args.workload, "particles": w["n"], "contact_points": w["c"], "grid": w["grid"],
                   "resample": "stratified, every step"},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": threads, "kind": kind, "sample": sample},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }))


def cpu_baseline(args, w):
    """The single-threaded CPU filter, as the reference runs it, on a bounded sample of the workload."""
    n_sample = min(w["n"], args.cpu_particles)
    value, dt, n = time_oracle(args.workload, n_sample, 3, 1, 1)
    kind = cpu_kind()
    return {"value": value, "unit": UNIT, "cores": 1, "kind": kind,
            "sample": "3 steps of %d of %d particles x %d contact points, 1 thread like the reference (host has %d cores); %s"
                      % (n, w["n"], w["c"], os.cpu_count(), KIND_TEXT[kind])}


def multi_gpu_parity(args, comm, rank, world, local_rank, n_local, patch, shard_mode, g_timed):
    """N > 1: the sharded filter against ONE GPU running the same global particle set (same seed, same steps), outside the
    timed region: CRC-32 of every rank's ancestor indices and weights after 3 steps must equal the CRC of the matching slice
    of the one-GPU run (rank 0 holds it; LOCAL_MIGRATE: one GPU with `world` logical shards).  Collective: every rank calls."""
    import zlib
    import numpy as np
    import torch.distributed as td
    from slam_eslam_b200 import GpuPoseEstimator, abi, synthetic
    g_timed.close()
    n_global = n_local * world
    kw = dict(min_effective=n_global + 1, patch_size=patch, shard_mode=shard_mode)
    g = GpuPoseEstimator(abi.default_config(**kw), device=local_rank, shard=comm.shard(n_local))
    comm.connect(g)
    contacts, step = synthetic.setup_filter(g, args.workload, n=n_local)
    for _ in range(3):
        r = g.stepFused(step, contacts)
    p = g.getParticles()
    mine = {"ancestors_crc": zlib.crc32(g.getAncestors().tobytes()), "weights_crc": zlib.crc32(p["weight"].tobytes()),
            "poses_crc": zlib.crc32(p["x"].tobytes() + p["y"].tobytes() + p["theta"].tobytes()), "ess": r.ess}
    g.close()
    allr = [None] * world
    td.all_gather_object(allr, mine)
    if rank != 0:
        return None
    one = GpuPoseEstimator(abi.default_config(virtual_shards=world if shard_mode == abi.SHARD_LOCAL_MIGRATE else 0, **kw), device=local_rank)
    synthetic.setup_filter(one, args.workload, n=n_global)
    for _ in range(3):
        r1 = one.stepFused(step, contacts)
    q = one.getParticles()
    anc = one.getAncestors()
    one.close()
    ok = True
    for k in range(world):
        sl = slice(k * n_local, (k + 1) * n_local)
        ref = {"ancestors_crc": zlib.crc32(anc[sl].tobytes()), "weights_crc": zlib.crc32(q["weight"][sl].tobytes()),
               "poses_crc": zlib.crc32(q["x"][sl].tobytes() + q["y"][sl].tobytes() + q["theta"][sl].tobytes()), "ess": r1.ess}
        ok = ok and ref == allr[k]
    return {"ancestors_crc": "%08x" % zlib.crc32(b"".join(("%08x" % a["ancestors_crc"]).encode() for a in allr)),
            "weights_crc": "%08x" % zlib.crc32(b"".join(("%08x" % a["weights_crc"]).encode() for a in allr)),
            "matches_1gpu": bool(ok), "steps": 3,
            "what": "CRC-32 per rank of ancestor indices, weights and poses after 3 steps == the same slices of one GPU running "
                    "all %d particles (bit-exact); ESS equal" % n_global}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=None,
                    help="timed steps; default 400 (~0.2 s on the GPU: enough nvidia-smi clock samples inside the timed region), "
                         "20 for --impl reference (a CPU step of the full workload takes ~0.2-0.5 s)")
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="C3")
    ap.add_argument("--cpu-particles", type=int, default=131072,
                    help="particles per step of the bounded CPU sample (the oracle does ~1.4e7 particle-updates/s)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--shard-mode", default="exact", choices=["exact", "local"],
                    help="N > 1: exact = exact-global resampling (ancestors identical to one GPU / the reference; the default and "
                         "the parity mode); local = LOCAL_MIGRATE (per-shard resampling, exact-global migration step when a shard collapses)")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.warmup < 3:
        args.warmup = 3
    if args.steps is None:
        args.steps = 20 if args.impl == "reference" else 400
    if args.impl == "reference":
        run_reference(args, rank, world)
        return

    import torch
    from slam_eslam_b200 import GpuPoseEstimator, abi, synthetic, dist
    w = synthetic.WORKLOADS[args.workload]
    n_local, c = w["n"], w["c"]
    if args.workload == "C5":                                            # 16M particles in total, sharded over the ranks
        n_local = (w["n"] // world) // 1024 * 1024
    torch.cuda.set_device(local_rank)
    comm = dist.init(world, rank, local_rank) if world > 1 else None
    n_global = n_local * world
    patch = 16 if args.workload == "C4" else 0                       # C4: EmbodiedSlamFilter per-particle map patches
    shard_mode = abi.SHARD_LOCAL_MIGRATE if args.shard_mode == "local" else abi.SHARD_EXACT_GLOBAL
    cfg = abi.default_config(min_effective=n_global + 1, patch_size=patch, shard_mode=shard_mode)   # resample every step (ESS <= N always)
    # Three passes over the SAME steps of the same trajectory (the same filter re-initialised with the same seed, W warm-up steps,
    # then K steps each): the synthetic filter converges -- its cloud contracts from 2 m to 0.1 m within ~50 resampling steps and
    # the weight kernel gets 10 % faster with it -- so legs measured one after the other without the reset would not see the
    # same work.
    #   1. value: K steps enqueued back to back, inputs resident, no event between the first and the last launch;
    #   2. the dominant kernel: the same K steps with CUDA events around it on every step (they cost the programmatic overlap at
    #      two places per step, ~4 us: not the pass the step time comes from);
    #   3. e2e: the same K steps (at most 100) through the reference-facing call with host inputs, the result read every step.
    g = GpuPoseEstimator(cfg, device=local_rank, shard=comm.shard(n_local) if comm else None)
    if comm:
        comm.connect(g)
    contacts, step = synthetic.setup_filter(g, args.workload, n=n_local)

    def restart():
        """The initial cloud again and the warm-up steps; sharded: every rank is idle before any rank re-initialises."""
        torch.cuda.synchronize()
        if comm: comm.barrier()
        synthetic.init_cloud(g, args.workload, n_local)
        if comm: comm.barrier()
        if args.warmup > 0: g.runSteps(step, contacts, args.warmup, kernel_events=False)

    # The clock sampler is started BEFORE the barrier (forking nvidia-smi takes ~20 ms on rank 0); run_steps itself begins with a
    # device-side rendezvous of all ranks, so no rank's events include waiting for a rank whose host got there later.
    if args.warmup > 0: g.runSteps(step, contacts, args.warmup, kernel_events=False)
    launches0 = g.launchCount()
    sampler = ClockSampler(local_rank) if rank == 0 else None
    if comm: comm.barrier()
    torch.cuda.synchronize()
    ms, _, res = g.runSteps(step, contacts, args.steps, kernel_events=False)
    torch.cuda.synchronize()
    if comm: comm.barrier()
    launches = g.launchCount() - launches0
    paths = g.pathCounters()

    restart()
    if comm: comm.barrier()
    torch.cuda.synchronize()
    ms_k, top_ms, _ = g.runSteps(step, contacts, args.steps)
    torch.cuda.synchronize()
    periods = sorted(g.stepTimes())
    median_ms = periods[len(periods) // 2] if periods else ms_k / args.steps
    if comm:
        ms = comm.max_float(ms)
        top_ms = comm.max_float(top_ms)
        median_ms = comm.max_float(median_ms)

    # end to end through the reference-facing call: host Step + contact array in, UpdateResult out, every step
    e2e_steps = max(3, min(args.steps, 100))
    restart()
    if comm: comm.barrier()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        r = g.stepFused(step, contacts)
    torch.cuda.synchronize()
    e2e_s = time.perf_counter() - t0
    if comm:
        e2e_s = comm.max_float(e2e_s)
    clocks = sampler.stop() if sampler else None
    h2d, d2h = g.ioBytes(len(contacts))
    parity = multi_gpu_parity(args, comm, rank, world, local_rank, n_local, patch, shard_mode, g) if comm else None
    if rank != 0:
        return

    updates = n_global * c
    value = updates * args.steps / (ms * 1e-3)
    lean = paths["lean"] > 0
    step_bytes, top_bytes = algorithmic_bytes(n_local, c, lean, patch)
    peak, peak_kind = measured_peak()
    top_avg_s = top_ms * 1e-3 / args.steps
    achieved = top_bytes / top_avg_s / 1e9
    top_kernel = "k_weight_lean" if lean else "k_project_weight"
    traffic = ncu_traffic(args.workload, top_kernel)
    out = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f64", "data": "synthetic",
        "config": {"workload": args.workload, "particles_per_gpu": n_local, "particles": n_global, "contact_points": c,
                   "grid": "%dx%d float2, replicated" % (w["grid"], w["grid"]), "resample": "stratified, every step", "patch_size": patch,
                   "sharding": ("none" if world == 1 else
                                "contiguous particle ranges, exact-global resampling (ancestors identical to 1 GPU)" if args.shard_mode == "exact" else
                                "contiguous particle ranges, LOCAL_MIGRATE: per-shard resampling, exact-global migration step when a shard collapses"),
                   "timing": "inputs larger than L2: particle state 2 x %.0f MB + scan buffers stream from HBM every "
                             "step; the %.0f MB map is L2-resident by design" % (n_local * 58 / 1e6, w["grid"] ** 2 * 8 / 1e6),
                   "step_algorithmic_bytes": step_bytes,
                   "step_hbm_frac": step_bytes / (ms * 1e-3 / args.steps) / 1e9 / peak,
                   "resampled_last_step": res.resampled, "ess_last_step": res.ess,
                   "lean_path_particles": paths["lean"], "generic_path_particles": paths["deferred"]},
        "median_ms_per_step": median_ms,
        "median_from": "the kernel-timing pass (events around the dominant kernel on every step: ~4 us per step above ms_per_step)",
        "roofline": {"bound": "hbm", "kernel": top_kernel, "achieved": achieved, "peak": peak,
                     "peak_kind": peak_kind + " (MEASURED_PEAKS.json hbm_gbs)" if peak_kind == "measured" else "fallback 6.65 TB/s",
                     "unit": "GB/s", "frac": achieved / peak, "traffic": traffic["bytes"] if traffic else None,
                     "traffic_source": traffic["source"] if traffic else None,
                     "achieved_kind": "algorithmic bytes per launch / measured launch time (CUDA events in this run)",
                     "limiter": "fp64 instruction issue" if lean else "fp64 instruction issue + L2 gathers",
                     "kernel_ms": top_avg_s * 1e3, "kernel_share_of_step": top_ms / ms,
                     "kernel_ms_sampling": "CUDA events around the kernel on every step of a second pass over the same steps (the filter re-initialised, same warm-up); ms_per_step comes from a pass without them; "
                                           "launch-bound sizes run as one CUDA graph: the events of its last replay",
                     "algorithmic_bytes_per_launch": top_bytes,
                     "note": "reported against the HBM roofline as the metric demands; the kernel itself is limited by instruction "
                             "issue (fp64 at half rate), not by HBM: the map is served from TMA-staged shared-memory tiles, so "
                             "DRAM traffic is far below the algorithmic bytes; see DESIGN.md §5 and profiles/"},
        "e2e": {"value": updates * e2e_steps / e2e_s, "unit": UNIT,
                "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h, "ms_per_step": 1e3 * e2e_s / e2e_steps},
        "gpu_launches": launches,
        "clocks": clocks,
    }
    if parity is not None:
        out["parity"] = parity
    if world == 1 and not args.no_cpu_baseline:
        out["cpu_baseline"] = cpu_baseline(args, w)
    emit(out)


if __name__ == "__main__":
    main()
