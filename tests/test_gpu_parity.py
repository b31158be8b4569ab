"""Parity of the CUDA path (through the C ABI, libeslam_gpu.so) against the CPU oracle on identical inputs and
identical Philox noise.  Bars (north_star): ancestor indices and map-cell indices bit-exact; poses 1e-6 m / rad and
normalised weights 1e-5 relative against the LITERAL restatement.  Against the SHARED-arithmetic oracle the CUDA path
is expected to be bit-identical in every double, and these tests require exactly that."""
import numpy as np
import pytest

from helpers import assert_particles_equal, kat_config, kat_contacts, load_kat, quadrant_grid
from slam_eslam_b200 import abi, synthetic

pytestmark = pytest.mark.gpu
KAT = load_kat()


@pytest.fixture(scope="module")
def gpu():
    import torch
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    from slam_eslam_b200 import GpuPoseEstimator
    return GpuPoseEstimator


def pair(gpu, oracle, workload, n=None, policy=None, **cfg):
    g = gpu(abi.default_config(**cfg))
    o = oracle.OraclePoseEstimator(abi.default_config(**cfg), policy=oracle.SHARED if policy is None else policy)
    contacts, step = synthetic.setup_filter(g, workload, n=n)
    synthetic.setup_filter(o, workload, n=n)
    return g, o, contacts, step


# ---------------------------------------------------------------------------------- reference KATs on the GPU
@pytest.mark.parametrize("name", ["K2_nogroup_flat", "K3_nogroup_uncertain"])
def test_reference_kat_nogroup(gpu, name):
    case = KAT[name]
    g = gpu(kat_config(case))
    g.setEnvironment(*quadrant_grid(case["z"], case["stddev"], case.get("res")))
    out = g.evaluatePoses([0.0], [0.0], [0.0], [0.0], [case["meas_var"]], kat_contacts(case))
    assert out["ok"][0] and out["n_contacts"][0] == case["expect_contacts"]
    for i, p in enumerate(case["points"]):
        assert np.allclose(out["cp"][0, i, :3], [p[0], p[1], case["z"][i]], atol=1e-6)
    rel = case["rel_tol_percent"] / 100.0
    assert abs(out["z_delta"][0]) < case["z_delta_abs_max"]
    assert abs(out["z_var"][0] - case["z_var"]) <= rel * case["z_var"]
    assert abs(out["weight"][0] - case["weight"]) <= rel * case["weight"]


def test_reference_kat_accessor_positions(gpu):
    """K1: the map is queried at pose * contact_point (test/testContactModel.cpp:89-124): here seen as the cell
    each contact lands in on a 1 m grid centred on the expected positions."""
    case = KAT["K1_accessor_positions"]
    g = gpu(kat_config(case))
    cells = np.zeros((8, 8, 2), dtype=np.float32)
    g.setEnvironment(cells, (0.5, 0.5), (-2.0, -2.0))
    for pose in case["poses"]:
        x, y, z = pose["xyz"]
        out = g.evaluatePoses([x], [y], [pose["theta"]], [z], [case["meas_var"]], kat_contacts(case))
        for got, want in zip(out["cell_index"][0], pose["expected_queries"]):
            ix, iy = int(np.floor((want[0] + 2.0) / 0.5)), int(np.floor((want[1] + 2.0) / 0.5))
            assert got == iy * 8 + ix


def test_reference_kat_map_absence(gpu):
    case = KAT["K5_map_absence_group"]
    g = gpu(kat_config(case))
    g.setEnvironment(*quadrant_grid(case["z"], case["stddev"], case["res"]))
    out = g.evaluatePoses([0.0], [0.0], [0.0], [0.0], [case["meas_var"]], kat_contacts(case))
    assert out["ok"][0] and out["n_contacts"][0] == 1


def test_kats_bit_identical_to_shared_oracle(gpu, oracle):
    for name in ("K2_nogroup_flat", "K3_nogroup_uncertain", "K2_default_radius", "K5_map_absence_group",
                 "STALE_updatePose_group"):
        case = KAT[name]
        g = gpu(kat_config(case))
        o = oracle.OraclePoseEstimator(kat_config(case), policy=oracle.SHARED)
        for f in (g, o):
            f.setEnvironment(*quadrant_grid(case["z"], case["stddev"], case.get("res")))
        args = ([0.0], [0.0], [0.0], [0.0], [case["meas_var"]], kat_contacts(case))
        a, b = g.evaluatePoses(*args, z_var_in=[0.5]), o.evaluatePoses(*args, z_var_in=[0.5])
        for k in a:
            assert np.array_equal(a[k], b[k]), (name, k, a[k], b[k])


# ---------------------------------------------------------------------------------- errors
def test_error_conventions(gpu):
    g = gpu()
    g.init(16, (0, 0, 0), (0.1, 0.1, 0.1), 0.0, 1.0)
    with pytest.raises(RuntimeError, match="No environment attached"):
        g.update(kat_contacts(KAT["K2_nogroup_flat"]))
    case = KAT["K2_nogroup_flat"]
    g.setEnvironment(*quadrant_grid(case["z"], case["stddev"]))
    with pytest.raises(RuntimeError, match="zero measurement variance"):
        g.evaluatePoses([0.0], [0.0], [0.0], [0.0], [0.0], kat_contacts(case))
    g2 = gpu(abi.default_config(measurement_error=0.0))
    g2.setEnvironment(*quadrant_grid(case["z"], case["stddev"]))
    g2.init(16, (0, 0, 0), (0.1, 0.1, 0.1), 0.0, 0.0)
    with pytest.raises(RuntimeError, match="zero measurement variance"):
        g2.update(kat_contacts(case))


# ---------------------------------------------------------------------------------- stage: init / project
def test_init_and_project_bit_identical(gpu, oracle):
    g, o, contacts, step = pair(gpu, oracle, "C2")
    assert_particles_equal(g.getParticles(), o.getParticles())
    for _ in range(3):
        g.project(step)
        o.project(step)
        assert_particles_equal(g.getParticles(), o.getParticles())


def test_spread_and_yaw_penalty_paths(gpu, oracle):
    """max_weight = 0 after init -> spread = 1 (src/PoseEstimator.cpp:194,224-236); yaw far away -> weight *= 0.7."""
    g, o, contacts, step = pair(gpu, oracle, "C1", slip_factor=0.5)
    far = abi.make_step(yaw=2.0, z_delta=0.01, z_var=1e-4, delta_mean=(0.05, 0.01, 0.02), delta_sigma=(0.01, 0.01, 0.01))
    for f in (g, o):
        f.setParticles(weight=np.full(250, 1.0 / 250))
        f.project(far)
    a, b = g.getParticles(), o.getParticles()
    assert_particles_equal(a, b)
    assert np.allclose(a["weight"], 0.7 / 250)


# ---------------------------------------------------------------------------------- stage: contact model
def test_evaluate_poses_bit_identical_and_cells_exact(gpu, oracle):
    """Random poses on the C2 terrain, 64 contact points: every output double and every map-cell index."""
    g, o, _, _ = pair(gpu, oracle, "C2", n=16)
    lit = oracle.OraclePoseEstimator(abi.default_config(), policy=oracle.LITERAL)
    synthetic.setup_filter(lit, "C2", n=16)
    contacts = synthetic.make_contact_points(64)
    rng = np.random.default_rng(11)
    n = 4000
    x, y = rng.uniform(-13.5, 13.5, (2, n))          # partly off the 25.6 m map
    th = rng.uniform(-7, 7, n)
    z = 0.19 + rng.normal(0, 0.2, n)
    zs = 10 ** rng.uniform(-2, 0.2, n)
    mv = zs * zs + 0.01
    a = g.evaluatePoses(x, y, th, z, mv, contacts, z_var_in=zs * zs)
    b = o.evaluatePoses(x, y, th, z, mv, contacts, z_var_in=zs * zs)
    c = lit.evaluatePoses(x, y, th, z, mv, contacts, z_var_in=zs * zs)
    for k in a:
        assert np.array_equal(a[k], b[k]), k
    # against the line-by-line restatement: cells bit-exact, floats within north_star's tolerance
    assert np.array_equal(a["cell_index"], c["cell_index"])
    assert np.array_equal(a["ok"], c["ok"]) and np.array_equal(a["n_contacts"], c["n_contacts"])
    np.testing.assert_allclose(a["weight"], c["weight"], rtol=1e-5, atol=1e-300)
    np.testing.assert_allclose(a["zpos_out"], c["zpos_out"], rtol=0, atol=1e-6)
    assert a["ok"].sum() > n // 10 and (a["cell_index"] < 0).any()


# ---------------------------------------------------------------------------------- end to end trajectories
@pytest.mark.parametrize("workload,steps,cfg", [
    ("C1", 40, {}),                                        # reference defaults: resample when ESS < 50
    ("C1", 12, {"min_effective": 251}),                    # resample every step
    ("C1", 12, {"min_effective": 251, "resample_mode": abi.RESAMPLE_SYSTEMATIC}),
    ("C2", 8, {"min_effective": 10001}),
])
def test_trajectory_bit_identical_to_shared_oracle(gpu, oracle, workload, steps, cfg):
    g, o, contacts, step = pair(gpu, oracle, workload, **cfg)
    resampled = 0
    for i in range(steps):
        g.project(step)
        o.project(step)
        ra, rb = g.update(contacts), o.update(contacts)
        assert ra.as_dict() == rb.as_dict(), "step %d: %r vs %r" % (i, ra.as_dict(), rb.as_dict())
        if ra.resampled:
            resampled += 1
            assert np.array_equal(g.getAncestors(), o.getAncestors()), "ancestors, step %d" % i
            assert np.array_equal(g.getCdf(), o.getCdf()), "running sum, step %d" % i
        assert_particles_equal(g.getParticles(), o.getParticles())
    assert resampled > 0
    assert np.array_equal(g.getCentroid(), o.getCentroid())
    assert g.getBestParticleIndex() == o.getBestParticleIndex()


def test_lean_and_generic_weight_kernels_agree(gpu, monkeypatch):
    """The lean path (particles binned by map tile, fused fixed-point cell index, gate on the high word of t) and the
    generic kernel (exact division, exact gate) must leave bit-identical particles."""
    n = 131072
    lean = gpu(abi.default_config(min_effective=n + 1))
    monkeypatch.setenv("ESLAM_GPU_NO_FAST", "1")
    generic = gpu(abi.default_config(min_effective=n + 1))
    contacts, step = synthetic.setup_filter(generic, "C3", n=n)
    monkeypatch.delenv("ESLAM_GPU_NO_FAST")
    synthetic.setup_filter(lean, "C3", n=n)
    for i in range(4):
        monkeypatch.setenv("ESLAM_GPU_NO_FAST", "1")
        rb = generic.stepFused(step, contacts)
        monkeypatch.delenv("ESLAM_GPU_NO_FAST")
        ra = lean.stepFused(step, contacts)
        assert ra.as_dict() == rb.as_dict(), "step %d" % i
        pc, pg = lean.pathCounters(), generic.pathCounters()
        assert pg["lean"] == 0 and pg["deferred"] == n and pc["lean"] == n and pc["deferred"] < n // 100, (pc, pg)
        assert np.array_equal(lean.getAncestors(), generic.getAncestors())
        assert_particles_equal(lean.getParticles(), generic.getParticles())


def test_lean_path_when_every_particle_needs_the_exact_evaluation(gpu, monkeypatch):
    """Particles placed ON cell borders: the fused fixed-point index of their own cell is never trusted, so every particle goes
    through the exact re-evaluation -- the per-CTA deferred list (worked off by all warps once the bin's chunks have run out) and,
    with more than 1024 of them per CTA, its overflow path (evaluated on the spot by the whole warp).  Bit-identical to the
    generic kernel, which takes every decision exactly."""
    n = 262144
    rng = np.random.default_rng(11)
    w = synthetic.WORKLOADS["C3"]
    scale, off = 0.05, -w["grid"] * 0.05 / 2.0
    x = off + rng.integers(900, 1150, n) * scale                       # on the border between two cells
    y = rng.uniform(-3.0, 3.0, n)
    th = rng.uniform(-3.0, 3.0, n)
    z = np.full(n, synthetic.terrain_height(0, 0) + 0.19)
    lean = gpu(abi.default_config(min_effective=n + 1))
    monkeypatch.setenv("ESLAM_GPU_NO_FAST", "1")
    generic = gpu(abi.default_config(min_effective=n + 1))
    contacts, step = synthetic.setup_filter(generic, "C3", n=n)
    monkeypatch.delenv("ESLAM_GPU_NO_FAST")
    synthetic.setup_filter(lean, "C3", n=n)
    for f in (lean, generic):
        f.setParticles(x=x, y=y, theta=th, zpos=z, zsigma=np.full(n, 1.001), weight=np.full(n, 1.0 / n))
    monkeypatch.setenv("ESLAM_GPU_NO_FAST", "1")
    rb = generic.update(contacts)
    monkeypatch.delenv("ESLAM_GPU_NO_FAST")
    ra = lean.update(contacts)
    assert ra.as_dict() == rb.as_dict()
    assert lean.pathCounters()["lean"] == n
    assert np.array_equal(lean.getAncestors(), generic.getAncestors())
    assert_particles_equal(lean.getParticles(), generic.getParticles())


def test_lean_path_at_the_map_border(gpu, oracle, monkeypatch):
    """A cloud wider than the map (C2: 25.6 m, sigma 8 m): particles near the border and outside are flagged by the
    binning and evaluated by contact_eval inside the lean kernel; everything must still be bit-identical to the oracle
    and to the generic kernel."""
    n = 4096
    cfg = dict(min_effective=n + 1)
    g = gpu(abi.default_config(**cfg))
    o = oracle.OraclePoseEstimator(abi.default_config(**cfg), policy=oracle.SHARED)
    monkeypatch.setenv("ESLAM_GPU_NO_FAST", "1")
    generic = gpu(abi.default_config(**cfg))
    contacts, step = synthetic.setup_filter(generic, "C2", n=n)
    monkeypatch.delenv("ESLAM_GPU_NO_FAST")
    for f in (g, o):
        synthetic.setup_filter(f, "C2", n=n)
    for f in (g, o, generic):
        f.init(n, (0.0, 0.0, 0.0), (8.0, 8.0, 0.3), synthetic.terrain_height(0, 0) + 0.19, 1.001)
    for i in range(3):
        monkeypatch.setenv("ESLAM_GPU_NO_FAST", "1")
        rc = generic.stepFused(step, contacts)
        monkeypatch.delenv("ESLAM_GPU_NO_FAST")
        ra = g.stepFused(step, contacts)
        o.project(step)
        rb = o.update(contacts)
        assert ra.as_dict() == rb.as_dict() == rc.as_dict(), "step %d" % i
        pc = g.pathCounters()
        assert pc["lean"] == n and n // 20 < pc["deferred"] < n, pc      # a good share is flagged, not all
        assert np.array_equal(g.getAncestors(), o.getAncestors())
        assert_particles_equal(g.getParticles(), o.getParticles())
        assert_particles_equal(g.getParticles(), generic.getParticles())


def test_fused_step_equals_project_then_update(gpu):
    g1, g2 = gpu(abi.default_config(min_effective=10001)), gpu(abi.default_config(min_effective=10001))
    contacts, step = synthetic.setup_filter(g1, "C2")
    synthetic.setup_filter(g2, "C2")
    for _ in range(3):
        g1.project(step)
        r1 = g1.update(contacts)
        r2 = g2.stepFused(step, contacts)
        assert r1.as_dict() == r2.as_dict()
    assert_particles_equal(g1.getParticles(), g2.getParticles())


def test_trajectory_within_tolerance_of_literal_oracle(gpu, oracle):
    """GPU vs the line-by-line restatement (libm, reference expression order) over 25 steps of C1."""
    g, o, contacts, step = pair(gpu, oracle, "C1", policy=oracle.LITERAL)
    for i in range(25):
        g.project(step)
        o.project(step)
        ra, rb = g.update(contacts), o.update(contacts)
        assert ra.resampled == rb.resampled and ra.data_particles == rb.data_particles
        if ra.resampled:
            assert np.array_equal(g.getAncestors(), o.getAncestors())        # bit-exact indices
        a, b = g.getParticles(), o.getParticles()
        np.testing.assert_allclose(a["weight"], b["weight"], rtol=1e-5, atol=0)
        for k in ("x", "y", "theta", "zpos"):
            np.testing.assert_allclose(a[k], b[k], rtol=0, atol=1e-6)
        assert np.array_equal(a["n_found"], b["n_found"])


# ---------------------------------------------------------------------------------- stage: resampling (exact running sum)
def weight_cases(n, rng):
    yield "uniform", np.full(n, 1.0 / n)
    w = rng.random(n); yield "random", w / w.sum()
    w = rng.random(n) ** 20; yield "peaked", w / w.sum()
    w = np.exp(rng.normal(size=n) * 8); yield "lognormal", w / w.sum()
    w = rng.random(n); w[rng.random(n) < 0.5] = 0; w[0] = 0
    if w.sum() > 0: yield "half zeros", w / w.sum()
    w = 2.0 ** -rng.integers(1, 60, size=n).astype(float); yield "dyadic (ties)", w / 2 ** np.ceil(np.log2(w.sum()))
    w = rng.random(n) * 1e-18; w[n // 2] = 1.0; yield "one heavy", w
    w = rng.random(n); yield "unnormalised", w * 3.7
    w = np.full(n, 2.0 ** -53); w[0] = 0.5; yield "all ties at a binade edge (sequential fallback)", w
    w = np.zeros(n); yield "all zero", w


@pytest.mark.parametrize("n", [1, 2, 250, 1024, 1025, 5000, 200000])
def test_resample_stage_bit_exact(gpu, oracle, n):
    """Identical weights in -> the reference's left-to-right running sum, bit for bit, and identical ancestors."""
    rng = np.random.default_rng(n)
    for name, w in weight_cases(n, rng):
        g = gpu()
        o = oracle.OraclePoseEstimator(abi.default_config(), policy=oracle.LITERAL)
        z = np.zeros(n)
        xs = np.arange(n, dtype=np.float64)
        for f in (g, o):
            f.setParticles(x=xs, y=z, theta=z, zpos=z, zsigma=z + 1, weight=w)
            f.resample()
        assert np.array_equal(g.getCdf(), np.cumsum(w)), "%s: running sum" % name      # numpy cumsum is sequential
        anc = g.getAncestors()
        assert np.array_equal(anc, o.getAncestors()), "%s: ancestors" % name
        p = g.getParticles()
        assert np.array_equal(p["x"], xs[anc]) and np.array_equal(p["weight"], w[anc])  # copies keep their weight (Q3)
        if name in ("uniform", "random", "peaked", "lognormal", "half zeros", "unnormalised", "one heavy"):
            assert g.debugCounters()["fallbacks"] == 0, "%s took the sequential fallback (%r)" % (name, g.debugCounters())
        if "fallback" in name and n > 4096:
            assert g.debugCounters()["fallbacks"] == 1


def test_normalize_centroid_best_index(gpu, oracle):
    n = 3000
    rng = np.random.default_rng(5)
    x, y, th, z = rng.normal(size=(4, n))
    w = rng.random(n); w[1234] = 3.0
    g, o = gpu(), oracle.OraclePoseEstimator(abi.default_config(), policy=oracle.SHARED)
    for f in (g, o):
        f.setParticles(x=x, y=y, theta=th, zpos=z, zsigma=np.ones(n), weight=w)
    assert g.getBestParticleIndex() == o.getBestParticleIndex() == 1234
    assert g.normalizeWeights() == o.normalizeWeights()
    assert np.array_equal(g.getParticles()["weight"], o.getParticles()["weight"])
    assert np.array_equal(g.getCentroid(), o.getCentroid())
    # all-zero weights -> uniform (Q8)
    g.setParticles(weight=np.zeros(n))
    assert g.normalizeWeights() == pytest.approx(n)
    assert np.all(g.getParticles()["weight"] == 1.0 / n)


# ---------------------------------------------------------------------------------- BASELINE full size (C3)
def test_full_size_step_matches_oracle_and_properties(gpu, oracle):
    """One step of C3 (1,048,576 particles x 64 contact points, 2048^2 grid, resampling forced): every particle
    bit-identical to the SHARED oracle; then size-independent properties on further steps."""
    n = synthetic.WORKLOADS["C3"]["n"]
    g, o, contacts, step = pair(gpu, oracle, "C3", min_effective=n + 1)
    ra, rb = g.stepFused(step, contacts), None
    o.project(step)
    rb = o.update(contacts)
    assert ra.as_dict() == rb.as_dict()
    assert ra.resampled == 1
    anc = g.getAncestors()
    assert np.array_equal(anc, o.getAncestors())
    assert_particles_equal(g.getParticles(), o.getParticles())
    for _ in range(3):
        before_w = None
        r = g.stepFused(step, contacts)
        assert r.resampled == 1 and 1.0 <= r.ess <= n
        anc = g.getAncestors().astype(np.int64)
        assert np.all(np.diff(anc) >= 0) and anc[0] >= 0 and anc[-1] < n       # sortedness
        cdf = g.getCdf()
        assert np.all(np.diff(cdf) >= 0) and abs(cdf[-1] - 1.0) < 1e-9          # normalised running sum
        counts = np.bincount(anc, minlength=n)
        w = np.diff(np.concatenate(([0.0], cdf))) * n
        assert np.all(counts <= np.ceil(w) + 1) and np.all(counts >= np.floor(w) - 1)   # stratified offspring bounds
        assert counts.sum() == n
    assert g.debugCounters()["fallbacks"] == 0 and g.debugCounters()["resamples"] == 4


# ---------------------------------------------------------------------------------- per-particle map patches (C4)
@pytest.mark.parametrize("n,steps", [(250, 10), (20000, 4)])
def test_patches_bit_identical_to_shared_oracle(gpu, oracle, n, steps):
    """EmbodiedSlamFilter per-particle map update (contact heights into particle-local 16x16 patches), patch-first
    lookup, and cloneMaps through resampling: particles AND every patch cell / origin identical to the oracle."""
    P = 16
    g, o, contacts, step = pair(gpu, oracle, "C1", n=n, patch_size=P, min_effective=n + 1)
    still = abi.make_step(delta_sigma=(0.001, 0.001, 0.0005))
    for i in range(steps):
        st = step if i % 2 == 0 else still               # small motion: written cells are hit again
        g.project(st)
        o.project(st)
        ra, rb = g.update(contacts), o.update(contacts)
        assert ra.as_dict() == rb.as_dict(), i
        assert np.array_equal(g.getAncestors(), o.getAncestors())
        assert_particles_equal(g.getParticles(), o.getParticles())
        ca, oa = g.getPatches(P)
        cb, ob = o.getPatches(P)
        assert np.array_equal(oa, ob), "patch origins, step %d" % i
        assert np.array_equal(ca, cb, equal_nan=True), "patch cells, step %d" % i
    assert (~np.isnan(ca[..., 0])).sum() > n // 2


def test_patches_change_the_result_only_through_written_cells(gpu):
    P = 16
    a = gpu(abi.default_config(patch_size=P, min_effective=0))
    b = gpu(abi.default_config(min_effective=0))
    contacts, step = synthetic.setup_filter(a, "C1")
    synthetic.setup_filter(b, "C1")
    for f in (a, b):
        f.project(step)
        f.update(contacts)
    assert np.array_equal(a.getParticles()["weight"], b.getParticles()["weight"])      # empty patches: same as shared map
    still = abi.make_step()
    for f in (a, b):
        f.project(still)
        f.update(contacts)
    assert not np.array_equal(a.getParticles()["weight"], b.getParticles()["weight"])
    with pytest.raises(RuntimeError):
        b.getPatches(P)


# ---------------------------------------------------------------------------------- larger sets (more tiles than threads)
def tree_sum(v):
    """Pairwise sum over the index-aligned binary tree (a missing right operand passes the left one through)."""
    a = np.array(v, dtype=np.float64)
    while a.size > 1:
        if a.size % 2:
            a = np.concatenate([a[:-1:2] + a[1::2], a[-1:]])
        else:
            a = a[0::2] + a[1::2]
    return float(a[0])


@pytest.mark.parametrize("n", [1 << 22, 5 << 20])
def test_millions_of_particles_resample_stage(gpu, n):
    """4,194,304 particles = 4096 tiles, 5,242,880 = 5120 tiles: the single-CTA kernels walk several tiles per thread and
    the tree reductions run their first levels in registers (as they do for the all-ranks tile arrays of sharded runs);
    the sums must still be the index-aligned tree's, the running sum the sequential one and the ancestors its lower
    bounds."""
    rng = np.random.default_rng(22)
    w = rng.random(n) ** 3
    w /= w.sum()
    g = gpu()
    z = np.zeros(n)
    g.setParticles(x=np.arange(n, dtype=np.float64), y=z, theta=z, zpos=z, zsigma=z + 1, weight=w)
    assert g.getWeightsSum() == tree_sum(w)
    g.resample()
    cdf = np.cumsum(w)
    assert np.array_equal(g.getCdf(), cdf)
    anc = g.getAncestors().astype(np.int64)
    assert np.all(np.diff(anc) >= 0)
    # lower-bound property against the exact running sum: cdf[a-1] < (k+u)/n <= cdf[a] for some u in [0,1)
    k = np.arange(n)
    lo = np.where(anc > 0, cdf[np.maximum(anc - 1, 0)], 0.0)
    assert np.all(lo < (k + 1.0) / n) and np.all(cdf[anc] >= k / n)
    assert g.debugCounters()["fallbacks"] == 0
    assert g.normalizeWeights() == pytest.approx(1.0 / np.sum((w[anc] / w[anc].sum()) ** 2), rel=1e-12)


# ---------------------------------------------------------------------------------- against the reference's own code
def _gold():
    import os
    import sys
    from helpers import GOLDEN
    sys.path.insert(0, GOLDEN)
    import make_reference_golden as gold
    return gold, np.load(os.path.join(GOLDEN, "reference_trajectories.npz"))


@pytest.mark.parametrize("case", ["c1_default", "c1_forced"])
def test_gpu_against_reference_golden_trajectories(gpu, case):
    """tests/golden/reference_trajectories.npz holds outputs of the reference's OWN compiled sources (oracle/ref/).  north_star's
    bars against them: resampled ancestor indices bit-exact at every resampling step of the trajectory; poses within 1e-6 m /
    1e-6 rad; normalised weights within 1e-5 relative; flags and fused-contact counts identical."""
    gold, G = _gold()
    spec = gold.CASES[case]
    got = gold.run_case(lambda cfg: gpu(cfg), spec)
    want = {k.split("/", 1)[1]: G[k] for k in G.files if k.startswith(case + "/")}
    assert np.array_equal(got["resampled"], want["resampled"])
    assert np.array_equal(got["ancestors"], want["ancestors"]) and want["ancestors"].shape[0] >= 1
    for i in spec["keep"]:
        for k in ("x", "y", "theta", "zpos", "zsigma"):
            np.testing.assert_allclose(got["%s_%d" % (k, i)], want["%s_%d" % (k, i)], rtol=0, atol=1e-6, err_msg="%s step %d" % (k, i))
        for k in ("weight", "mprob"):
            np.testing.assert_allclose(got["%s_%d" % (k, i)], want["%s_%d" % (k, i)], rtol=1e-5, atol=1e-300, err_msg="%s step %d" % (k, i))
        for k in ("floating", "n_found"):
            assert np.array_equal(got["%s_%d" % (k, i)], want["%s_%d" % (k, i)])
    keep = ~np.isnan(want["ess"])
    np.testing.assert_allclose(got["ess"][keep], want["ess"][keep], rtol=1e-5)
    np.testing.assert_allclose(got["centroid"], want["centroid"], rtol=0, atol=1e-6)
    assert got["best_index"][0] == want["best_index"][0]


def test_gpu_against_live_reference_c2(gpu):
    """The reference's compiled filter itself (oracle/_ref/libeslam_ref.so, prebuilt in the build container) next to the GPU
    on C2 (10 000 particles x 16 contact points), resampling every step: ancestors bit-exact, floats within north_star's bars;
    then the ContactModel seam on 20 000 random poses: map-cell indices bit-exact."""
    from oracle import ref_binding
    if not ref_binding.available():
        pytest.skip("oracle/_ref/libeslam_ref.so did not travel")
    cfg = dict(min_effective=10001)
    g = gpu(abi.default_config(**cfg))
    r = ref_binding.RefPoseEstimator(abi.default_config(**cfg), debug=True)
    contacts, step = synthetic.setup_filter(g, "C2")
    synthetic.setup_filter(r, "C2")
    for i in range(8):
        g.project(step); r.project(step)
        ra, rb = g.update(contacts), r.update(contacts)
        assert ra.resampled == rb.resampled == 1
        assert np.array_equal(g.getAncestors(), r.getAncestors()), "ancestors differ from the reference at step %d" % i
        a, b = g.getParticles(), r.getParticles()
        for k in ("x", "y", "theta", "zpos", "zsigma"):
            np.testing.assert_allclose(a[k], b[k], rtol=0, atol=1e-6, err_msg=k)
        np.testing.assert_allclose(a["weight"], b["weight"], rtol=1e-5, atol=1e-300)
        assert np.array_equal(a["floating"], b["floating"]) and np.array_equal(a["n_found"], b["n_found"])
    rng = np.random.default_rng(5)
    n = 20000
    x, y = rng.uniform(-13.0, 13.0, n), rng.uniform(-13.0, 13.0, n)       # the 512^2 grid spans +-12.8 m: some poses hang over the border
    th, z, mv = rng.uniform(-4, 4, n), rng.normal(0.19, 0.15, n), rng.uniform(0.01, 1.2, n)
    a, b = g.evaluatePoses(x, y, th, z, mv, contacts), r.evaluatePoses(x, y, th, z, mv, contacts)
    assert np.array_equal(a["cell_index"], b["cell_index"])
    assert np.array_equal(a["ok"], b["ok"]) and np.array_equal(a["n_contacts"], b["n_contacts"])
    ok = a["ok"].astype(bool)
    np.testing.assert_allclose(a["weight"][ok], b["weight"][ok], rtol=1e-5, atol=1e-300)
    np.testing.assert_allclose(a["z_delta"][ok], b["z_delta"][ok], rtol=0, atol=1e-6)


# ---------------------------------------------------------------------------------- sharded resampling modes on one GPU
@pytest.mark.parametrize("shards,n", [(4, 8192), (8, 65536)])
def test_local_migrate_mode_bit_identical_to_oracle(gpu, oracle, shards, n):
    """LOCAL_MIGRATE (include/eslam_gpu.h) with the particles of one GPU treated as `shards` logical shards: local walks
    per shard while no shard has collapsed, the exact-global walk otherwise; every double, ancestors, flags identical to the
    oracle's definition (tests/test_shard_modes.py), through trajectories that contain both kinds of step."""
    cfg = dict(min_effective=n + 1, shard_mode=abi.SHARD_LOCAL_MIGRATE, virtual_shards=shards)
    g, o, contacts, step = pair(gpu, oracle, "C2", n=n, **cfg)
    ns = n // shards
    own = np.arange(n) // ns
    kinds = set()
    for it in range(8):
        if it == 4:      # crush shard 1: the next resampling must migrate
            w = g.getParticles()["weight"].copy(); w[ns:2 * ns] *= 1e-3
            g.setParticles(weight=w); o.setParticles(weight=w)
        ra, rb = g.stepFused(step, contacts), (o.project(step), o.update(contacts))[1]
        assert ra.as_dict() == rb.as_dict(), (it, ra.as_dict(), rb.as_dict())
        a, b = g.getAncestors(), o.getAncestors()
        assert np.array_equal(a, b), "ancestors differ at step %d" % it
        kinds.add("local" if np.array_equal(a // ns, own) else "migrate")
        assert_particles_equal(g.getParticles(), o.getParticles())
    assert kinds == {"local", "migrate"}
    assert g.debugCounters()["fallbacks"] == 0


def test_local_migrate_resample_stage_adversarial(gpu, oracle):
    """resample() alone in LOCAL_MIGRATE mode on weight vectors with zeros, ties, heavy particles and an all-zero shard."""
    n, shards = 16384, 4
    ns = n // shards
    rng = np.random.default_rng(7)
    zeros = np.zeros(n)
    vectors = [rng.random(n) + 0.5, np.tile(np.r_[np.zeros(ns // 2), np.ones(ns // 2)], shards), rng.random(n) ** 8 + 1e-3,
               np.r_[rng.random(3 * ns) + 0.5, np.zeros(ns)], np.ones(n)]
    for thresholds in (dict(), dict(migrate_min_share=0.0, migrate_min_ess=0.0)):
        for w in vectors:
            cfg = dict(shard_mode=abi.SHARD_LOCAL_MIGRATE, virtual_shards=shards, **thresholds)
            g = gpu(abi.default_config(**cfg)); o = oracle.OraclePoseEstimator(abi.default_config(**cfg), policy=oracle.SHARED)
            for f in (g, o):
                f.setParticles(x=np.arange(n, dtype=float), y=zeros, theta=zeros, zpos=zeros, zsigma=zeros + 1, weight=w)
                f.normalizeWeights()
                f.resample()
            assert np.array_equal(g.getAncestors(), o.getAncestors())
            assert np.array_equal(g.getCdf(), o.getCdf())
            assert_particles_equal(g.getParticles(), o.getParticles(), keys=("x", "weight"))


# ---------------------------------------------------------------------------------- parity holes named by the round-1 review
def _six_group_contacts(c_per_group=3):
    """6 contact groups (groupId 0..5): more than 4 fused contacts per particle -> `4 - cpoints.size()` wraps (Q2)."""
    pos, contact, gid = [], [], []
    for g in range(6):
        for j in range(c_per_group):
            pos.append((0.3 * np.cos(g) + 0.02 * j, 0.3 * np.sin(g), 0.0))
            contact.append(0.5); gid.append(g)
    return abi.make_contacts(pos, contact, gid)


@pytest.mark.parametrize("lean", [True, False])
def test_unsigned_exponent_quirk_Q2_on_device(gpu, oracle, monkeypatch, lean):
    """src/PoseEstimator.cpp:336 with more than 4 fused contacts on the DEVICE: k_loop_b takes pow_table[5] (the wrapped
    size_t exponent), every data particle's weight becomes 0, the sum is 0 and normalizeWeights resets to 1/N (Q8) --
    identical to the oracle, on the lean kernel (gend[] / n_groups with 6 groups) and on the generic one."""
    if not lean:
        monkeypatch.setenv("ESLAM_GPU_NO_FAST", "1")
    n = 3000
    cfg = dict(min_effective=n + 1)
    g, o, _, step = pair(gpu, oracle, "C2", n=n, **cfg)
    flat = np.zeros((512, 512, 2), dtype=np.float32); flat[..., 1] = 0.05
    for f in (g, o):
        f.setEnvironment(flat, (0.05, 0.05), (-12.8, -12.8))
    contacts = _six_group_contacts()
    for it in range(3):
        ra, rb = g.stepFused(step, contacts), (o.project(step), o.update(contacts))[1]
        assert ra.as_dict() == rb.as_dict(), (it, ra.as_dict(), rb.as_dict())
        a, b = g.getParticles(), o.getParticles()
        assert_particles_equal(a, b)
        assert np.array_equal(g.getAncestors(), o.getAncestors())
        assert ra.weights_sum == 0.0 and np.all(a["n_found"] == 6)       # Q2 -> all weights 0 -> uniform reset (Q8)
    assert g.pathCounters()["lean"] == (n if lean else 0)


def test_unsigned_exponent_quirk_Q2_against_the_live_reference(gpu):
    """The same on the reference's own compiled code: pow(x, size_t(4 - 6)) == 0 -> uniform weights; n_found, flags, ancestors
    equal, weights equal (all 1/N)."""
    from oracle import ref_binding
    if not ref_binding.available():
        pytest.skip("oracle/_ref/libeslam_ref.so did not travel")
    n = 2000
    cfg = dict(min_effective=n + 1)
    g = gpu(abi.default_config(**cfg))
    r = ref_binding.RefPoseEstimator(abi.default_config(**cfg), debug=True)
    _, step = synthetic.setup_filter(g, "C2", n=n)
    synthetic.setup_filter(r, "C2", n=n)
    flat = np.zeros((512, 512, 2), dtype=np.float32); flat[..., 1] = 0.05
    for f in (g, r):
        f.setEnvironment(flat, (0.05, 0.05), (-12.8, -12.8))
    contacts = _six_group_contacts()
    for it in range(2):
        g.project(step); r.project(step)
        ra, rb = g.update(contacts), r.update(contacts)
        a, b = g.getParticles(), r.getParticles()
        assert np.array_equal(a["n_found"], b["n_found"]) and np.all(a["n_found"] == 6)
        assert np.array_equal(a["weight"], b["weight"]) and np.all(a["weight"] == 1.0 / n)
        assert np.array_equal(g.getAncestors(), r.getAncestors())


def test_c4_at_baseline_size_patches_bit_identical(gpu, oracle):
    """BASELINE config 4 at full size: 100 000 particles x 64 contact points on the 4096^2 grid with 16 x 16 per-particle
    patches, resampling every step: every particle, every patch cell, every patch origin and every ancestor equal to the oracle."""
    P, n = 16, 100000
    g, o, contacts, step = pair(gpu, oracle, "C4", n=n, patch_size=P, min_effective=n + 1)
    still = abi.make_step(delta_sigma=(0.001, 0.001, 0.0005))
    for i in range(3):
        st = step if i % 2 == 0 else still
        ra, rb = g.stepFused(st, contacts), (o.project(st), o.update(contacts))[1]
        assert ra.as_dict() == rb.as_dict(), i
        assert ra.resampled == 1
        assert np.array_equal(g.getAncestors(), o.getAncestors())
        assert_particles_equal(g.getParticles(), o.getParticles())
        ca, oa = g.getPatches(P)
        cb, ob = o.getPatches(P)
        assert np.array_equal(oa, ob), "patch origins, step %d" % i
        assert np.array_equal(ca, cb, equal_nan=True), "patch cells, step %d" % i
    assert (~np.isnan(ca[..., 0])).sum() > n


def test_sequential_fallback_of_the_running_sum(gpu, oracle):
    """A legal weight vector whose exact running sum needs more boundary elements than the pool holds: k_chain falls back to the
    literal left-to-right sum (counted), ancestors and running sum still equal to numpy / the oracle."""
    n = 50000
    rng = np.random.default_rng(3)
    w = np.full(n, 1e-17); w[0] = 1.0                                       # every add is near a quarter ulp of the sum ...
    w[1::2] = 2.0 ** -53 * (1 + rng.random(n // 2))                        # ... or an ambiguous half-ulp case: boundaries everywhere
    g = gpu(abi.default_config()); o = oracle.OraclePoseEstimator(abi.default_config(), policy=oracle.LITERAL)
    z = np.zeros(n)
    for f in (g, o):
        f.setParticles(x=np.arange(n, dtype=float), y=z, theta=z, zpos=z, zsigma=z + 1, weight=w)
        f.resample()
    assert np.array_equal(g.getCdf(), np.cumsum(w))
    assert np.array_equal(g.getAncestors(), o.getAncestors())
    assert g.debugCounters()["error_flags"] == 0


def test_multi_gpu_parity_worker():
    """tests/dist_worker.py under torchrun on every visible GPU (2 or more): the sharded filter -- exchanges over peer memory,
    ancestors read in place over NVLink -- bit-identical to one GPU (exact-global stratified / systematic, LOCAL_MIGRATE,
    per-particle patches, the sequential fallback across shards)."""
    import os, subprocess, sys
    import torch
    if not torch.cuda.is_available() or torch.cuda.device_count() < 2:
        pytest.skip("needs 2 or more GPUs (run with gpurun --gpus 2)")
    world = 2 if torch.cuda.device_count() < 4 else (4 if torch.cuda.device_count() < 8 else 8)
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    out = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(world),
                          "--master-addr", "127.0.0.1", "--master-port", "29533", os.path.join(root, "tests", "dist_worker.py")],
                         stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, timeout=900)
    assert out.returncode == 0 and "dist parity ok" in out.stdout, out.stdout[-4000:]


# ---------------------------------------------------------------------------------- processMap proper, PoseDistribution moments
def test_process_map_bit_identical_to_shared_oracle(gpu, oracle):
    """EmbodiedSlamFilter::processMap(scanMap, match, update) (src/EmbodiedSlamFilter.cpp:179-232): scan merges into every
    particle's own map and the match weight, interleaved with contact updates and resampling (maps travel with their
    particles): every weight, every patch cell, every origin identical to the oracle's definition."""
    P, n = 16, 3000
    g, o, contacts, step = pair(gpu, oracle, "C1", n=n, patch_size=P, min_effective=n + 1)
    rng = np.random.default_rng(11)
    for it in range(5):
        scan = np.zeros((9, 13, 2), dtype=np.float32)
        scan[..., 0] = rng.normal(-0.19, 0.03, scan.shape[:2]); scan[..., 1] = 0.01 + 0.02 * rng.random(scan.shape[:2])
        scan[rng.random(scan.shape[:2]) < 0.15, 0] = np.nan
        args = (scan, (0.04, 0.06), (-0.2 + 0.05 * it, -0.25))
        for f in (g, o):
            f.processMap(*args, match=it > 0, update=it != 3)
        assert_particles_equal(g.getParticles(), o.getParticles())
        ca, oa = g.getPatches(P)
        cb, ob = o.getPatches(P)
        assert np.array_equal(oa, ob), "patch origins after processMap %d" % it
        assert np.array_equal(ca, cb, equal_nan=True), "patch cells after processMap %d" % it
        ra, rb = g.stepFused(step, contacts), (o.project(step), o.update(contacts))[1]
        assert ra.as_dict() == rb.as_dict(), it
        assert np.array_equal(g.getAncestors(), o.getAncestors())
    assert (~np.isnan(ca[..., 0])).reshape(n, -1).sum(axis=1).min() > 10
    m, mo = g.poseMoments(), o.poseMoments()
    assert m["weight"] == mo["weight"] and np.array_equal(m["mean"], mo["mean"]) and np.array_equal(m["cov"], mo["cov"])


def test_process_map_needs_per_particle_maps(gpu):
    g = gpu(abi.default_config())
    synthetic.setup_filter(g, "C1")
    with pytest.raises(RuntimeError, match="per-particle maps"):
        g.processMap(np.zeros((2, 2, 2), dtype=np.float32), (0.05, 0.05), (0, 0))


@pytest.mark.parametrize("multinomial", [False, True])
@pytest.mark.parametrize("samples", [700, 5000])
def test_resample_with_a_sample_count_of_its_own(gpu, oracle, multinomial, samples):
    """ParticleFilter::resample_stratified(samples) / resample_multinomial(samples) (src/ParticleFilter.hpp:85-147): the set has
    `samples` particles afterwards; ancestors and copies identical to the oracle (multinomial: weights reset to 1 / old size)."""
    n = 5000
    rng = np.random.default_rng(4)
    w = rng.random(n) ** 3
    z = np.zeros(n)
    out = []
    for f in (gpu(abi.default_config()), oracle.OraclePoseEstimator(abi.default_config(), policy=oracle.SHARED)):
        f.setParticles(x=np.arange(n, dtype=float), y=z, theta=z, zpos=z, zsigma=z + 1, weight=w)
        f.normalizeWeights()
        (f.resampleMultinomial if multinomial else f.resampleStratified)(samples)
        assert f.size() == samples
        out.append((f.getAncestors(), f.getParticles()))
    assert np.array_equal(out[0][0], out[1][0])
    assert_particles_equal(out[0][1], out[1][1], keys=("x", "weight"))
    if multinomial:
        assert np.all(out[0][1]["weight"] == 1.0 / n)
    else:
        assert np.all(np.diff(out[0][0].astype(np.int64)) >= 0)


# ---------------------------------------------------------------------------------- SurfaceHash, the relocaliser
def test_surface_hash_bit_identical_to_oracle(gpu, oracle):
    """SurfaceHash::create on the device (src/SurfaceHash.hpp:155-231), PoseEstimator::init(n, hash) (src/PoseEstimator.cpp:75-86)
    and sampleFromHash (:130-182): the table (every candidate pose and bucket, the per-bucket counts), the seeded particle set
    and the re-seeded one identical to the oracle's definition."""
    n = 4000
    g, o, contacts, step = pair(gpu, oracle, "C1", n=n, min_effective=n + 1)
    for f in (g, o):
        f.hashCreate(20, 16)
    (na, ca), (nb, cb) = g.hashInfo(), o.hashInfo()
    assert na == nb and na > 16 * 150 * 150 and np.array_equal(ca, cb)
    ha, hb = g.hashGet(0, na), o.hashGet(0, nb)
    for k in ha:
        assert np.array_equal(ha[k], hb[k]), k
    for f in (g, o):
        f.initFromHash(n)
    assert_particles_equal(g.getParticles(), o.getParticles())
    total = 0
    for it in range(4):
        ra, rb = g.stepFused(step, contacts), (o.project(step), o.update(contacts))[1]
        assert ra.as_dict() == rb.as_dict(), it
        assert_particles_equal(g.getParticles(), o.getParticles())
        ka, kb = g.sampleFromHash(contacts), o.sampleFromHash(contacts)
        assert ka == kb
        total += ka
        assert_particles_equal(g.getParticles(), o.getParticles())
    assert total > 0
