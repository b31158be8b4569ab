"""Pins the CPU oracle against the REFERENCE'S OWN CODE.

oracle/_ref/libeslam_ref.so is /root/reference/src/{ContactModel,PoseEstimator}.cpp and the headers they include
(ParticleFilter.hpp, PoseParticle.hpp, Configuration.hpp, ...), compiled unmodified against stand-in headers for the absent
dependencies (oracle/ref_shim, oracle/ref/ref_driver.cpp), its noise served from the same Philox counters.

* tests against tests/golden/reference_trajectories.npz (outputs of that library, committed with the script that made
  them) run everywhere: the LITERAL oracle must reproduce every stored double, flag and ancestor index BIT FOR BIT;
* tests marked `live` also run the library itself (built here from /root/reference, or the prebuilt file that travelled to
  the GPU box) on further seeds / configurations, on the ContactModel seam, on the reference's known-answer tests and on
  the reference's own unit-test binary.
"""
import os
import re
import subprocess
import sys

import numpy as np
import pytest

from helpers import GOLDEN, assert_particles_equal, kat_config, kat_contacts, load_kat
from slam_eslam_b200 import abi, synthetic

sys.path.insert(0, GOLDEN)
import make_reference_golden as gold  # noqa: E402

from oracle import ref_binding  # noqa: E402

live = pytest.mark.skipif(not ref_binding.available(), reason="oracle/_ref/libeslam_ref.so not built and /root/reference absent")
GOLD = np.load(os.path.join(GOLDEN, "reference_trajectories.npz"))
KAT = load_kat()


@pytest.mark.parametrize("case", sorted(gold.CASES))
def test_literal_oracle_reproduces_the_reference_bit_for_bit(oracle, case):
    spec = gold.CASES[case]
    got = gold.run_case(lambda cfg: oracle.OraclePoseEstimator(cfg, policy=oracle.LITERAL), spec)
    want = {k.split("/", 1)[1]: GOLD[k] for k in GOLD.files if k.startswith(case + "/")}
    assert set(got) == set(want)
    assert np.array_equal(got["resampled"], want["resampled"])
    assert int(want["resampled"].sum()) >= 1
    for k in sorted(want):
        if k in ("ess", "centroid"):
            continue
        assert np.array_equal(got[k], want[k]), "%s/%s differs from the reference" % (case, k)
    keep = ~np.isnan(want["ess"])                       # update() does not return the ESS of a step that resampled
    assert np.array_equal(got["ess"][keep], want["ess"][keep])
    assert np.array_equal(got["centroid"][[0, 1, 3]], want["centroid"][[0, 1, 3]])     # x, y, z: same sums, same order
    assert abs(got["centroid"][2] - want["centroid"][2]) < 1e-12                        # heading: read back from a rotation


def _pair(oracle, **cfg):
    r = ref_binding.RefPoseEstimator(abi.default_config(**cfg), debug=True)
    o = oracle.OraclePoseEstimator(abi.default_config(**cfg), policy=oracle.LITERAL)
    return r, o


@live
@pytest.mark.parametrize("cfg", [
    dict(seed=7, min_effective=120),
    dict(seed=2**40 + 3, min_effective=251, slip_factor=0.5, max_yaw_deviation=0.05),
    dict(seed=11, min_effective=251, spread_threshold=2.0, discount_factor=0.8, min_contacts=2, measurement_error=0.02),
    dict(seed=5, min_effective=251, use_shape_update=0, contact_likelihood_correction=1.0, contact_point_radius=0.0),
])
def test_live_reference_trajectory_bit_exact(oracle, cfg):
    """Whole filter steps on further seeds and configurations: slip, yaw penalty, spreading (max_weight carried between
    steps), minContacts, discounting -- every field of every particle, ancestors, ESS, flags identical."""
    r, o = _pair(oracle, **cfg)
    contacts, step = synthetic.setup_filter(r, "C1")
    synthetic.setup_filter(o, "C1")
    step = abi.make_step(yaw=0.1, z_delta=0.003, z_var=2e-6, delta_mean=(0.02, 0.001, 0.005), delta_sigma=(0.004, 0.002, 0.003))
    assert_particles_equal(r.getParticles(), o.getParticles(), keys=("x", "y", "theta", "zpos", "zsigma", "weight"))
    resamples = 0
    for i in range(25):
        r.project(step); o.project(step)
        assert_particles_equal(r.getParticles(), o.getParticles(), keys=("x", "y", "theta", "zpos", "zsigma", "weight"))
        ra, rb = r.update(contacts), o.update(contacts)
        assert ra.resampled == rb.resampled, i
        assert_particles_equal(r.getParticles(), o.getParticles())
        if rb.resampled:
            resamples += 1
            assert np.array_equal(r.getAncestors(), o.getAncestors())
        else:
            assert ra.ess == rb.ess
            assert (ra.data_particles, ra.total_points) == (rb.data_particles, rb.total_points)
    assert resamples >= 1
    assert r.getBestParticleIndex() == o.getBestParticleIndex()
    ca, cb = r.getCentroid(), o.getCentroid()
    assert np.array_equal(ca[[0, 1, 3]], cb[[0, 1, 3]]) and abs(ca[2] - cb[2]) < 1e-12


@live
def test_live_reference_contact_model_seam_bit_exact(oracle):
    """ContactModel::evaluatePose / evaluateWeight / updateZPositionEstimate on 2000 random poses over the synthetic
    terrain (incl. poses over the map border and over empty cells): map-cell indices and every output double identical."""
    r, o = _pair(oracle)
    cells, scale, offset = synthetic.make_terrain(200, 200)
    r.setEnvironment(cells, scale, offset); o.setEnvironment(cells, scale, offset)
    contacts = synthetic.make_contact_points(20)
    rng = np.random.default_rng(3)
    n = 2000
    x, y = rng.uniform(-5.6, 5.6, n), rng.uniform(-5.6, 5.6, n)
    th, z = rng.uniform(-4, 4, n), rng.normal(0.19, 0.15, n)
    mv, zv = rng.uniform(0.01, 1.2, n), rng.uniform(0.001, 1.1, n)
    a = r.evaluatePoses(x, y, th, z, mv, contacts, z_var_in=zv)
    b = o.evaluatePoses(x, y, th, z, mv, contacts, z_var_in=zv)
    assert 0.2 < a["ok"].mean() < 1.0 and (a["cell_index"] < 0).any()
    assert np.array_equal(a["cell_index"], b["cell_index"])
    assert np.array_equal(a["ok"], b["ok"]) and np.array_equal(a["n_contacts"], b["n_contacts"])
    ok = a["ok"].astype(bool)
    for k in ("weight", "z_delta", "z_var", "zpos_out", "z_var_out"):
        assert np.array_equal(a[k][ok], b[k][ok]), k
    assert np.array_equal(a["pose_var"], b["pose_var"])
    for i in range(n):
        m = a["n_contacts"][i]
        assert np.array_equal(a["cp"][i, :m], b["cp"][i, :m])


@live
def test_live_reference_resample_adversarial_weights(oracle):
    """ParticleFilter::resample on weight vectors with ties, zeros, one heavy particle and tiny tails."""
    rng = np.random.default_rng(0)
    n = 4096
    vectors = [rng.random(n), np.r_[np.zeros(n // 2), np.ones(n // 2)], np.r_[1.0, np.full(n - 1, 1e-12)],
               rng.random(n) ** 20, np.r_[np.full(n - 1, 1e-300), 1.0]]
    for w in vectors:
        r, o = _pair(oracle)
        zeros = np.zeros(n)
        for f in (r, o):
            f.setParticles(x=np.arange(n, dtype=float), y=zeros, theta=zeros, zpos=zeros, zsigma=zeros + 1, weight=w)
            f.normalizeWeights()
            f.resample()
        assert np.array_equal(r.getAncestors(), o.getAncestors())
        assert_particles_equal(r.getParticles(), o.getParticles(), keys=("x", "weight"))


@live
def test_live_reference_known_answer_tests():
    """The known-answer vectors of tests/golden/contact_model_kat.json through the reference's compiled ContactModel."""
    def make(case):
        f = ref_binding.RefPoseEstimator(kat_config(case))
        f.setFakeMap(case["z"], case["stddev"], case.get("res"))
        return f
    try:
        case = KAT["K1_accessor_positions"]
        for pose in case["poses"]:
            f = make(case)
            x, y, z = pose["xyz"]
            f.evaluatePoses([x], [y], [pose["theta"]], [z], [case["meas_var"]], kat_contacts(case))
            q = f.queriedPoints()
            for got, want in zip(q, pose["expected_queries"]):
                assert np.linalg.norm(got - np.array(want)) < case["tol"]
        for name in ("K2_nogroup_flat", "K3_nogroup_uncertain"):
            case = KAT[name]
            out = make(case).evaluatePoses([0.0], [0.0], [0.0], [0.0], [case["meas_var"]], kat_contacts(case))
            assert bool(out["ok"][0]) is case["expect_ok"] and out["n_contacts"][0] == case["expect_contacts"]
            assert abs(out["z_delta"][0]) < case["z_delta_abs_max"]
            rel = case["rel_tol_percent"] / 100.0
            assert abs(out["z_var"][0] - case["z_var"]) <= rel * case["z_var"]
            assert abs(out["weight"][0] - case["weight"]) <= rel * case["weight"]
        case = KAT["K2_default_radius"]
        out = make(case).evaluatePoses([0.0], [0.0], [0.0], [0.0], [case["meas_var"]], kat_contacts(case))
        assert out["ok"][0] and out["n_contacts"][0] == 4 and abs(out["z_delta"][0] - case["z_delta"]) < 1e-12
    finally:
        ref_binding.load().eslam_ref_clear_fake_map()


@live
def test_live_reference_lowest_point_heuristic(oracle):
    for groups in ([-1, -1, -1, -1], [0, 0, 1, 1], [0, 0, 0, 1]):
        pts = abi.make_contacts([(-1, -1, 0.1), (1, -1, -0.1), (-1, 1, 0.1), (1, 1, -0.1)], [1, 2, 3, 4], groups)
        for upd in (False, True):
            a, pa = ref_binding.lowest_points(pts, upd)
            b, pb = oracle.lowest_points(pts, upd)
            assert np.array_equal(a, b) and np.array_equal(pa, pb)


@pytest.mark.skipif(not (os.path.exists(ref_binding.TEST_BIN) or ref_binding.have_sources()), reason="needs /root/reference or the prebuilt binary")
def test_the_references_own_unit_test_runs_against_its_own_code():
    """test/testContactModel.cpp + src/ContactModel.cpp, both unmodified, with a stand-in Boost.Test.  What passes and what
    fails is what SURVEY.md §8c derived by hand (the failing checks are stale against the reference's LIVE code: they assume
    contactPointRadius == 0 and minContacts <= 2), so the exact outcome is asserted."""
    if not os.path.exists(ref_binding.TEST_BIN):
        ref_binding.build()
    p = subprocess.run([ref_binding.TEST_BIN], capture_output=True, text=True)
    cases = {m.group(1): (int(m.group(2)), int(m.group(3))) for m in re.finditer(r"CASE (\w+) checks=(\d+) failures=(\d+)", p.stdout)}
    assert cases["test_lowest_points_without_groups"] == (13, 0)
    assert cases["test_lowest_points_with_group"] == (11, 0)
    # K2/K3: contact selection, zVar 0.5 / 2 and weight 1 hold; only |zDelta| < 1e-6 fails (= +0.01, the default radius)
    assert cases["test_updatePose_nogroup"] == (26, 2)
    assert len(re.findall(r"FAILED .*getZDelta", p.stdout)) == 2
    # K1: the four norm checks are off by exactly the default contactPointRadius
    assert cases["test_eslam_passes_valid_global_position_to_map_accessor"] == (4, 4)
    assert len(re.findall(r"\[0\.01 != 1e-06\]", p.stdout)) == 6
    # two groups give two fused contacts < minContacts (3): evaluatePose returns false, BOOST_REQUIRE stops the case
    assert cases["test_updatePose_group"] == (1, 1)
    assert cases["test_mapAbsence_group"] == (1, 1)


@live
@pytest.mark.parametrize("seed", range(12))
def test_live_reference_fuzz_configurations(oracle, seed):
    """Random configurations, contact sets (grouped / ungrouped / unknown and sub-threshold contact probabilities, up to 64
    points), odometry streams, maps with many empty cells and a translated map frame (C_global2local): the compiled reference
    and the LITERAL oracle stay bit-identical through every step, ancestors included."""
    rng = np.random.default_rng(1000 + seed)
    n = int(rng.choice([64, 250, 1000]))
    cfg = dict(seed=int(rng.integers(1, 2**31)), min_effective=int(rng.choice([n // 4, n // 2, n + 1])),
               measurement_error=float(rng.choice([0.05, 0.1, 0.3])), discount_factor=float(rng.choice([0.8, 0.9, 1.0])),
               spread_threshold=float(rng.choice([0.0, 0.9, 1.5])), slip_factor=float(rng.choice([0.0, 0.05, 0.6])),
               max_yaw_deviation=float(rng.choice([0.0, 0.05, 0.26])), min_contacts=int(rng.choice([1, 2, 3, 4])),
               contact_likelihood_correction=float(rng.choice([0.33, 1.0])), contact_point_radius=float(rng.choice([0.0, 0.01, 0.05])),
               use_shape_update=int(rng.choice([0, 1])))
    # map: 120 x 90 cells of 0.05-0.1 m, 15 % empty, offset so that part of the cloud hangs over the border
    w, h = 120, 90
    sx, sy = float(rng.choice([0.05, 0.1])), float(rng.choice([0.05, 0.08]))
    cells = np.empty((h, w, 2), dtype=np.float32)
    cells[..., 0] = (0.1 * rng.standard_normal((h, w))).astype(np.float32)
    cells[..., 1] = (0.01 + 0.05 * rng.random((h, w))).astype(np.float32)
    cells[..., 0][rng.random((h, w)) < 0.15] = np.nan
    offset = (-w * sx / 2 + 0.3, -h * sy / 2 - 0.2)
    g2l = None
    if seed % 2:
        g2l = np.array([[1, 0, 0, rng.uniform(-0.5, 0.5)], [0, 1, 0, rng.uniform(-0.5, 0.5)], [0, 0, 1, 0.0]])
    # contacts: 4..64 points; grouped runs, some ungrouped, unknown (NaN) and sub-threshold probabilities
    c = int(rng.choice([4, 8, 20, 64]))
    pos = np.c_[rng.uniform(-0.4, 0.4, c), rng.uniform(-0.4, 0.4, c), rng.uniform(-0.2, 0.05, c)]
    prob = rng.choice([0.5, 0.1, np.nan, 1.0, 0.2], size=c, p=[0.6, 0.1, 0.1, 0.1, 0.1])
    if seed % 3 == 0:
        gid = np.full(c, -1)
    else:
        gid = np.sort(rng.integers(0, 4, c))                # contiguous groups 0..3 (Q2: at most four fused contacts)
    contacts = abi.make_contacts(pos, prob, gid)
    r, o = _pair(oracle, **cfg)
    for f in (r, o):
        f.setEnvironment(cells, (sx, sy), offset, g2l=g2l)
        f.init(n, (0.1, -0.2, 0.3), (0.8, 0.6, 0.2), 0.15, 0.5)
    assert_particles_equal(r.getParticles(), o.getParticles(), keys=("x", "y", "theta", "zpos", "zsigma", "weight"))
    touched = 0
    for i in range(12):
        step = abi.make_step(yaw=0.3 + 0.01 * i, z_delta=float(rng.normal(0, 0.002)), z_var=float(rng.choice([0.0, 2e-6, 1e-4])),
                             delta_mean=(0.03, float(rng.normal(0, 0.005)), 0.01), delta_sigma=(0.005, 0.004, 0.002))
        r.project(step); o.project(step)
        ra, rb = r.update(contacts), o.update(contacts)
        assert ra.resampled == rb.resampled, (i, cfg)
        assert_particles_equal(r.getParticles(), o.getParticles())
        touched += int(rb.data_particles)
        if rb.resampled:
            assert np.array_equal(r.getAncestors(), o.getAncestors()), (i, cfg)
        else:
            assert ra.ess == rb.ess and (ra.data_particles, ra.total_points) == (rb.data_particles, rb.total_points)
    FUZZ_TOUCHED.append(touched)


FUZZ_TOUCHED = []


@live
def test_live_reference_fuzz_was_not_vacuous():
    """Most fuzz configurations must have had particles in contact with the map (a run where every particle floats compares
    nothing of the contact model)."""
    if len(FUZZ_TOUCHED) < 12:
        pytest.skip("runs after the twelve fuzz cases of this module")
    assert sum(1 for t in FUZZ_TOUCHED if t > 0) >= 8, FUZZ_TOUCHED


@pytest.mark.skipif(not ref_binding.have_sources(), reason="needs /root/reference")
def test_the_references_particle_filter_unit_test_walks_off_the_end():
    """test/UnitTest.cpp (ParticleFilter<State> on a toy tracker) resamples unnormalised weights; resample_stratified has no
    bounds guard (src/ParticleFilter.hpp:99-103), so with weights summing below a stratum's draw it reads past the end of
    xi_k.  Built with AddressSanitizer the reference's own test stops exactly there -- the behaviour SURVEY.md §8 (a14) flags,
    and the reason the oracle and the kernels clamp the ancestor to the last particle instead of reproducing the read."""
    out = os.path.join(os.path.dirname(ref_binding.LIB_PATH), "unit_test_asan")
    r = subprocess.run(["make", "-s", "-C", os.path.join(os.path.dirname(ref_binding.LIB_PATH), "..", "ref"),
                        "REF=" + ref_binding.REFERENCE, "../_ref/unit_test_asan"], capture_output=True, text=True)
    if r.returncode != 0:
        pytest.skip("AddressSanitizer build not available: " + r.stderr[-200:])
    p = subprocess.run([out], capture_output=True, text=True, env=dict(os.environ, ASAN_OPTIONS="detect_leaks=0"))
    assert p.returncode != 0
    assert "heap-buffer-overflow" in p.stderr
    assert re.search(r"resample_stratified.*ParticleFilter\.hpp:10[0-3]", p.stderr)
