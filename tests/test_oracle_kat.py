"""Pin the CPU oracle against the reference's own known-answer tests (test/testContactModel.cpp), both
arithmetic policies, through the reference's FakeMLSAccess and through the equivalent 2x2 dense grid."""
import math

import numpy as np
import pytest

from helpers import kat_config, kat_contacts, load_kat, quadrant_grid

KAT = load_kat()
POLICIES = [0, 1]  # LITERAL, SHARED


def make(oracle, case, policy, use_grid):
    f = oracle.OraclePoseEstimator(kat_config(case), policy=policy)
    if use_grid:
        f.setEnvironment(*quadrant_grid(case["z"], case["stddev"], case.get("res")))
    else:
        f.setFakeMap(case["z"], case["stddev"], case.get("res"))
    return f


@pytest.mark.parametrize("policy", POLICIES)
def test_K1_accessor_receives_world_positions(oracle, policy):
    case = KAT["K1_accessor_positions"]
    for pose in case["poses"]:
        f = make(oracle, case, policy, use_grid=False)
        x, y, z = pose["xyz"]
        f.evaluatePoses([x], [y], [pose["theta"]], [z], [case["meas_var"]], kat_contacts(case))
        q = f.queriedPoints()
        assert len(q) >= 2, "No mls points have been accessed."
        for got, want in zip(q, pose["expected_queries"]):
            assert np.linalg.norm(got - np.array(want)) < case["tol"]


@pytest.mark.parametrize("use_grid", [False, True])
@pytest.mark.parametrize("policy", POLICIES)
@pytest.mark.parametrize("name", ["K2_nogroup_flat", "K3_nogroup_uncertain"])
def test_K2_K3_update_pose_nogroup(oracle, name, policy, use_grid):
    case = KAT[name]
    f = make(oracle, case, policy, use_grid)
    out = f.evaluatePoses([0.0], [0.0], [0.0], [0.0], [case["meas_var"]], kat_contacts(case))
    assert bool(out["ok"][0]) is case["expect_ok"]
    assert out["n_contacts"][0] == case["expect_contacts"]
    # check_contact_point_selection (test/testContactModel.cpp:53-67): point = (x, y, z[index])
    for i, p in enumerate(case["points"]):
        assert np.allclose(out["cp"][0, i, :3], [p[0], p[1], case["z"][i]], atol=1e-6)
    assert abs(out["z_delta"][0]) < case["z_delta_abs_max"]                       # BOOST_CHECK_SMALL
    rel = case["rel_tol_percent"] / 100.0                                         # BOOST_CHECK_CLOSE is in percent
    assert abs(out["z_var"][0] - case["z_var"]) <= rel * case["z_var"]
    assert abs(out["weight"][0] - case["weight"]) <= rel * case["weight"]


@pytest.mark.parametrize("policy", POLICIES)
def test_K2_default_radius(oracle, policy):
    case = KAT["K2_default_radius"]
    f = make(oracle, case, policy, use_grid=False)
    out = f.evaluatePoses([0.0], [0.0], [0.0], [0.0], [case["meas_var"]], kat_contacts(case))
    assert out["ok"][0] and out["n_contacts"][0] == 4
    assert abs(out["z_delta"][0] - case["z_delta"]) < 1e-12
    assert abs(out["z_var"][0] - case["z_var"]) < 1e-12
    assert abs(out["weight"][0] - 1.0) < 1e-12
    assert abs(out["pose_var"][0] - case["pose_var"]) < 1e-12


@pytest.mark.parametrize("name", ["K4_lowest_without_groups", "K4_lowest_with_groups"])
def test_K4_lowest_point_heuristic(oracle, name):
    case = KAT[name]
    contacts = kat_contacts(case)
    lowest, probs = oracle.lowest_points(contacts, update_probabilities=False)
    assert len(lowest) == len(case["expected_lowest"])
    for got, idx in zip(lowest, case["expected_lowest"]):
        assert list(got) == case["points"][idx]
    assert list(probs) == case["contact"]                 # probabilities untouched
    _, probs = oracle.lowest_points(contacts, update_probabilities=True)
    assert list(probs) == case["expected_contact_after_update"]


@pytest.mark.parametrize("use_grid", [False, True])
@pytest.mark.parametrize("policy", POLICIES)
def test_K5_map_absence_drops_group(oracle, policy, use_grid):
    case = KAT["K5_map_absence_group"]
    f = make(oracle, case, policy, use_grid)
    out = f.evaluatePoses([0.0], [0.0], [0.0], [0.0], [case["meas_var"]], kat_contacts(case))
    assert out["ok"][0]
    assert out["n_contacts"][0] == case["expect_contacts"]


@pytest.mark.parametrize("policy", POLICIES)
def test_stale_group_test_live_behaviour(oracle, policy):
    case = KAT["STALE_updatePose_group"]
    f = make(oracle, case, policy, use_grid=False)
    out = f.evaluatePoses([0.0], [0.0], [0.0], [0.0], [case["meas_var"]], kat_contacts(case))
    assert not out["ok"][0]
    assert out["n_contacts"][0] == case["expect_contacts"]
    for i, p in enumerate(case["first_point_of_group"]):
        assert np.allclose(out["cp"][0, i, :3], p, atol=1e-12)


def test_zero_measurement_variance_is_an_error(oracle):
    case = KAT["K2_nogroup_flat"]
    f = make(oracle, case, 0, use_grid=False)
    with pytest.raises(RuntimeError, match="zero measurement variance"):
        f.evaluatePoses([0.0], [0.0], [0.0], [0.0], [0.0], kat_contacts(case))


def test_no_environment_is_an_error(oracle):
    f = oracle.OraclePoseEstimator()
    f.init(4, (0, 0, 0), (0.1, 0.1, 0.1), 0.0, 1.0)
    with pytest.raises(RuntimeError, match="No environment attached"):
        f.update(kat_contacts(KAT["K2_nogroup_flat"]))


def test_sticky_group_valid_quirk_Q4(oracle):
    """If the FIRST evaluated point of a group misses the map, group_valid is never reset
    (src/ContactModel.cpp:189-190 vs :210-213) and every following group is skipped."""
    case = dict(KAT["K5_map_absence_group"])
    case["res"] = [False, True, True, True]      # first point (-1,-1) -> index 0 misses
    f = make(oracle, case, 0, use_grid=False)
    out = f.evaluatePoses([0.0], [0.0], [0.0], [0.0], [1.0], kat_contacts(case))
    assert out["n_contacts"][0] == 0 and not out["ok"][0]
    assert len(f.queriedPoints()) == 1           # nothing after the first miss is even looked up
