"""Pin oracle/ransac.py (the OpenGV / GTSAM / libstdc++ restatement).

 * RNG: against this image's real libstdc++ (oracle/rng_check.cpp compiled with g++).
 * 2-pt / 5-pt / 3-pt / 1-pt: replay of the reference's synthetic-scene tests
   (tests/testTracker.cpp:704-801, 804-895, 898-1039, 1042-1185): the inlier / outlier sets must be
   classified exactly and the translation recovered within the reference's tolerances.
"""
import os
import shutil
import subprocess

import numpy as np
import pytest

import scenes
from kimera_vio_b200.params import CameraParams
from oracle import ransac as rs
from oracle.rig import StereoRig

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.skipif(shutil.which("g++") is None, reason="needs g++")
def test_rnd_table_matches_libstdcxx(tmp_path):
    exe = str(tmp_path / "rng_check")
    subprocess.check_call(["g++", "-O2", "-std=c++17", "-o", exe, os.path.join(ROOT, "oracle", "rng_check.cpp")])
    want = np.array(subprocess.check_output([exe, "rnd", "3000"]).split(), np.int64)
    assert np.array_equal(want, rs.rnd_table(3000, 12345, "lemire"))
    legacy = rs.rnd_table(3000, 12345, "legacy")
    assert legacy.min() >= 0 and legacy.max() < 2 ** 31 and not np.array_equal(legacy, want)


def _rig():
    return StereoRig(CameraParams.euroc_left(), CameraParams.euroc_right())


# tests/testTracker.cpp:804-895
@pytest.mark.parametrize("planar,n_in,n_out", [(False, 80, 0), (False, 80, 20), (True, 80, 20)])
def test_2pt_given_rotation_exact_sets(planar, n_in, n_out):
    cam = CameraParams.euroc_left()
    R, T = np.eye(3), np.array([1.0, 0, 0])
    for rep in range(3):
        rng = np.random.default_rng(3 + rep)
        f_ref, f_cur = scenes.mono_scene(rng, cam, R, T, n_in, n_out, planar)
        prob = rs.Problem2d2dGivenRot(f_ref, f_cur, R, rs.rnd_table(4096))
        ok, pose, inl = rs.run_ransac(prob, 1e-6, 100, 0.995)
        assert ok
        assert inl == list(range(n_in))
        t = pose[:, 3]
        assert np.allclose(t, T / np.linalg.norm(T), atol=1e-3)


# tests/testTracker.cpp:704-801
@pytest.mark.parametrize("planar,n_in,n_out", [(False, 82, 0), (False, 80, 40), (True, 80, 40)])
def test_5pt_nister_exact_sets(planar, n_in, n_out):
    cam = CameraParams.euroc_left()
    R, T = scenes.expmap([0.01, 0.01, 0.01]), np.array([1.0, 0, 0])
    rng = np.random.default_rng(3)
    f_ref, f_cur = scenes.mono_scene(rng, cam, R, T, n_in, n_out, planar)
    prob = rs.Problem2d2dNister(f_ref, f_cur, rs.rnd_table(16384))
    ok, pose, inl = rs.run_ransac(prob, 1e-6, 1000, 0.995)
    assert ok
    assert inl == list(range(n_in))
    assert np.allclose(pose[:, :3], R, atol=1e-3)
    assert np.allclose(pose[:, 3] / np.linalg.norm(pose[:, 3]), T, atol=1e-3)


# tests/testTracker.cpp:898-1039
@pytest.mark.parametrize("n_in,n_out", [(3, 0), (40, 0), (80, 40)])
def test_3pt_arun_exact_sets(n_in, n_out):
    rig = _rig()
    R, T = scenes.expmap([0.1, 0.1, 0.1]), np.array([rig.baseline, 0, 0])
    rng = np.random.default_rng(3)
    sc = scenes.stereo_scene(rng, rig, R, T, n_in, n_out, [rig.baseline * 10, rig.baseline * 20])
    prob = rs.Problem3d3d(sc["p_ref"], sc["p_cur"], rs.rnd_table(4096))
    ok, pose, inl = rs.run_ransac(prob, 0.3, 100, 0.995)
    assert ok
    assert inl == list(range(n_in))
    if n_out == 0:
        assert np.allclose(pose[:, 3], T, atol=1e-3)
    for i in range(n_in):
        exp = pose[:, :3].T @ sc["p_ref"][i] - pose[:, :3].T @ pose[:, 3]
        assert np.linalg.norm(exp - sc["p_cur"][i]) < (1e-3 if n_out == 0 else 1e-1)


# tests/testTracker.cpp:1042-1185
@pytest.mark.parametrize("n_in,n_out", [(3, 0), (40, 0), (80, 40)])
def test_1pt_voting_exact_sets(n_in, n_out):
    rig = _rig()
    R, T = scenes.expmap([0.1, 0.1, 0.1]), np.array([rig.baseline, 0, 0])
    rng = np.random.default_rng(3)
    sc = scenes.stereo_scene(rng, rig, R, T, n_in, n_out, [rig.baseline * 10, rig.baseline * 20])
    # the reference test stores the rectified-frame points as keypoints_3d for this solver
    calib = (rig.fx, rig.fy, rig.cx, rig.cy, rig.baseline)
    R1 = rig.R1
    p_ref = (R1 @ sc["p_ref"].T).T
    p_cur = (R1 @ sc["p_cur"].T).T
    Rrect = R1 @ R @ R1.T
    matches = [(i, i) for i in range(n_in + n_out)]
    status, pose, inl, info = rs.outlier_rejection_3d3d_given_rotation(
        sc["ref_left"], sc["ref_right"], sc["cur_left"], sc["cur_right"], p_ref, p_cur, calib,
        matches, Rrect, 1.0, 5)  # TrackerParams default ransac_threshold_stereo
    assert inl == list(range(n_in))
    assert status == (rs.VALID if n_in >= 5 else rs.FEW_MATCHES)
    assert np.allclose(pose[:, 3], R1 @ T, atol=1e-2)
    assert np.all(np.linalg.eigvalsh(info) > 0)


def test_backproject2_jacobian_numeric():
    fx, fy, cx, cy, b = 436.2, 436.2, 364.4, 256.9, 0.11
    z0 = (400.0, 380.0, 240.0)
    p, J = rs.backproject2_jacobian(*z0, fx, fy, cx, cy, b)
    Jn = np.zeros((3, 3))
    for k in range(3):
        d = [0.0, 0.0, 0.0]
        d[k] = 1e-5
        pp, _ = rs.backproject2_jacobian(z0[0] + d[0], z0[1] + d[1], z0[2] + d[2], fx, fy, cx, cy, b)
        pm, _ = rs.backproject2_jacobian(z0[0] - d[0], z0[1] - d[1], z0[2] - d[2], fx, fy, cx, cy, b)
        Jn[:, k] = (pp - pm) / 2e-5
    assert np.allclose(J, Jn, rtol=1e-5, atol=1e-6)


def test_voting_mahalanobis_formula_equivalence():
    """The closed-form f32 Mahalanobis distance used by the 1-point voting (Tracker.cpp:495-520) equals
    v' O^-1 v on random SPD matrices (tests/testTracker.cpp:1480-1528, tolerance 1e-2): a pair (i, j)
    votes for each other exactly when that distance is below the threshold."""
    rng = np.random.default_rng(12)
    for _ in range(200):
        m = rng.uniform(-1, 1, (3, 5))
        O = (m @ m.T).astype(np.float32)
        v = rng.uniform(-1, 1, 3).astype(np.float32)
        ref = float(v.astype(np.float64) @ np.linalg.solve(O.astype(np.float64), v.astype(np.float64)))
        # two relative translations v and 0 with covariances O and 0: one pair, distance = v' O^-1 v
        rel = np.stack([v, np.zeros(3, np.float32)])
        cov = np.stack([O, np.zeros((3, 3), np.float32)])
        above = rs.voting_1pt(rel, cov, np.float32(ref * (1 - 1e-3) - 1e-2))
        below = rs.voting_1pt(rel, cov, np.float32(ref * (1 + 1e-3) + 1e-2))
        assert above[0] == 1 and below[0] == 2, (ref, above[0], below[0])


# ----------------------------------------------------------------------------------------------
# The reference's tests on the reference's OWN seeded scenes (tests/golden/tracker_scenes.npz, built by
# tests/golden/make_tracker_scenes.py from glibc srand(3)/rand() and libstdc++ normal_distribution streams)
# ----------------------------------------------------------------------------------------------
SCENES = os.path.join(ROOT, "tests", "golden", "tracker_scenes.npz")


def seeded_scenes():
    import json
    z = np.load(SCENES)
    cams = json.loads(str(z["cams_json"]))

    def cam(d):
        d = dict(d)
        d["T_BS"] = np.asarray(d["T_BS"], np.float64)
        return CameraParams(**d)
    return z, cam(cams["left"]), cam(cams["right"])


@pytest.mark.parametrize("order", ["gcc", "clang"])
def test_reference_seeded_scenes(order):
    """What tests/testTracker.cpp asserts -- every synthesized inlier kept, every outlier removed, VALID status,
    translation / point tolerances -- must hold for the oracle on the scenes the reference's RNG streams produce;
    and the oracle's inlier lists, iteration counts and number of RNG draws must equal the recorded ones."""
    z, left, right = seeded_scenes()
    rig = StereoRig(left, right)
    R5, Rs = z["R_5pt"], z["R_stereo"]
    # geometricOutlierRejection2d2d, testTracker.cpp:704-801 (ransac_max_iterations 1000)
    for ci in range(3):
        pre = "%s/5pt/%d/" % (order, ci)
        planar, n_in, n_out = [int(v) for v in z[pre + "meta"]]
        prob = rs.Problem2d2dNister(z[pre + "f_ref"], z[pre + "f_cur"], rs.rnd_table(16384))
        ok, model, inl, its = rs.sac_ransac(prob, 1e-6, 1000, 0.995)
        assert ok and inl == list(range(n_in))                       # :784-797
        assert inl == [int(v) for v in z[pre + "oracle_inliers"]]
        assert its == int(z[pre + "oracle_iterations"]) and prob._rnd_pos == int(z[pre + "oracle_draws"])
        assert np.allclose(model, z[pre + "oracle_pose"], atol=1e-9)
    # geometricOutlierRejection2d2dGivenRotation, testTracker.cpp:804-895
    for ci in range(3):
        pre = "%s/2pt/%d/" % (order, ci)
        planar, n_in, n_out = [int(v) for v in z[pre + "meta"]]
        prob = rs.Problem2d2dGivenRot(z[pre + "f_ref"], z[pre + "f_cur"], np.eye(3), rs.rnd_table(4096))
        ok, model, inl, its = rs.sac_ransac(prob, 1e-6, 100, 0.995)
        assert ok and inl == list(range(n_in))                       # :879-891
        assert its == int(z[pre + "oracle_iterations"]) and prob._rnd_pos == int(z[pre + "oracle_draws"])
        assert np.allclose(model[:, 3], [1.0, 0, 0], atol=1e-3)
    # geometricOutlierRejection3d3d, testTracker.cpp:898-1039 (ransac_threshold_stereo 0.3)
    T = np.array([rig.baseline, 0, 0])
    for ci in range(4):
        pre = "%s/3pt/%d/" % (order, ci)
        planar, n_in, n_out = [int(v) for v in z[pre + "meta"]]
        p_ref, p_cur = z[pre + "p_ref"], z[pre + "p_cur"]
        prob = rs.Problem3d3d(p_ref, p_cur, rs.rnd_table(4096))
        ok, pose, inl = rs.run_ransac(prob, 0.3, 100, 0.995)
        assert ok and inl == list(range(n_in))                       # :985-1014
        assert inl == [int(v) for v in z[pre + "oracle_inliers"]]
        tol = 1e-3 if ci < 2 else 1e-1
        if ci < 2:
            assert np.allclose(pose[:, 3], T, atol=tol)              # :1019-1024
        for i in range(n_in):                                        # :1029-1036
            exp = pose[:, :3].T @ p_ref[i] - pose[:, :3].T @ pose[:, 3]
            assert np.linalg.norm(exp - p_cur[i]) < tol
    # geometricOutlierRejection3d3dGivenRotation, testTracker.cpp:1042-1185
    calib = (rig.fx, rig.fy, rig.cx, rig.cy, rig.baseline)
    for ci in range(4):
        pre = "%s/1pt/%d/" % (order, ci)
        planar, n_in, n_out = [int(v) for v in z[pre + "meta"]]
        p_ref = (rig.R1 @ z[pre + "p_ref"].T).T
        p_cur = (rig.R1 @ z[pre + "p_cur"].T).T
        n = len(p_ref)
        status, pose, inl, info = rs.outlier_rejection_3d3d_given_rotation(
            z[pre + "ref_left"], z[pre + "ref_right"], z[pre + "cur_left"], z[pre + "cur_right"], p_ref, p_cur, calib,
            [(i, i) for i in range(n)], rig.R1 @ Rs @ rig.R1.T, 1.0, 5)
        assert inl == list(range(n_in))                              # :1133-1161
        assert inl == [int(v) for v in z[pre + "oracle_inliers"]] and status == int(z[pre + "oracle_status"])
