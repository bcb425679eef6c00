"""Stage-level exports that complete the drop-in boundary (include/kvfe.h, "the remaining public methods of the
replaced classes"): each against the oracle's restatement of the reference method on real and synthetic inputs."""
import cv2
import numpy as np
import pytest

import helpers as H
from kimera_vio_b200.params import CameraParams, FrontendParams
from oracle import frontend as ofe
from oracle import ransac as ors
from oracle.rig import StereoRig

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def env():
    p, rig, ctx = H.euroc_setup(batch=2)          # batch 2: the stage calls must park the second stream
    orig = StereoRig(CameraParams.euroc_left(), CameraParams.euroc_right())
    yield dict(p=p, rig=rig, ctx=ctx, orig=orig)
    ctx.close()


def _points(n=400, seed=5):
    rng = np.random.default_rng(seed)
    pts = np.stack([rng.uniform(-5, 757, n), rng.uniform(-5, 485, n)], 1).astype(np.float32)
    edge = np.array([[0, 0], [751, 479], [0.4, 479.6], [751.5, 0.49], [375.5, 239.5], [-0.5, 100], [100, 479.51]], np.float32)
    return np.concatenate([pts, edge])


def test_undistort_rectify_left_and_check(env):
    o, ctx = env["orig"], env["ctx"]
    kps = _points()
    kps = kps[(kps[:, 0] >= 0) & (kps[:, 0] <= 751) & (kps[:, 1] >= 0) & (kps[:, 1] <= 479)]
    exp = ofe.undistort_rectify_left_keypoints([tuple(q) for q in kps], o)
    es = np.array([s for s, _ in exp], np.int32)
    ex = np.array([q for _, q in exp], np.float32)
    st, xy = ctx.undistort_rectify_left_keypoints(kps)
    assert np.array_equal(st, es) and np.array_equal(xy, ex)
    # the check alone, with the caller's undistorted points and other tolerances
    und = ofe.undistort_rectify_keypoints([tuple(q) for q in kps], o.left, o.R1, o.P1)
    for tol in (2.0, 0.25, 10.0):
        exp = ofe.undistort_rectify_left_keypoints([tuple(q) for q in kps], o, pixel_tol=tol)
        st, xy = ctx.check_rectified_keypoints(0, kps, np.array(und, np.float32), tol)
        assert np.array_equal(st, np.array([s for s, _ in exp], np.int32))
        assert np.array_equal(xy, np.array([q for _, q in exp], np.float32))
    assert (es != 0).any() and (es == 0).any()


def test_distort_unrectify(env):
    o, ctx = env["orig"], env["ctx"]
    rng = np.random.default_rng(9)
    n = 300
    xy = np.stack([rng.uniform(0, 751, n), rng.uniform(0, 479, n)], 1).astype(np.float32)
    st = rng.integers(0, 4, n).astype(np.int32)
    exp = np.array(ofe.distort_unrectify_right_keypoints([(int(s), (q[0], q[1])) for s, q in zip(st, xy)], o), np.float32)
    got = ctx.distort_unrectify_keypoints(1, st, xy)
    assert np.array_equal(got, exp)


def test_right_keypoints_and_depth(env):
    p, o, ctx = env["p"], env["orig"], env["ctx"]
    m = ofe.StereoMatcher(p, o)
    g, lefts, rights = H.golden()
    L, R = lefts[0], rights[0]
    sf = ofe.StereoFrame.make(0, 0, L, R, o)
    c = cv2.goodFeaturesToTrack(L, 250, 0.001, 20).reshape(-1, 2).astype(np.float32)
    lk = ofe.undistort_rectify_left_keypoints([tuple(q) for q in c], o)
    ls = np.array([s for s, _ in lk], np.int32)
    lx = np.array([q for _, q in lk], np.float32)
    # rectified images as the reference holds them in the StereoFrame
    Lr = cv2.remap(L, o.map_lx, o.map_ly, cv2.INTER_LINEAR)
    Rr = cv2.remap(R, o.map_rx, o.map_ry, cv2.INTER_LINEAR)
    exp = m.get_right_keypoints_rectified(Lr, Rr, lk, o.fx, o.baseline)
    ers = np.array([s for s, _ in exp], np.int32)
    erx = np.array([q for _, q in exp], np.float32)
    rs, rx = ctx.right_keypoints_rectified(Lr, Rr, ls, lx)
    assert np.array_equal(rs, ers)
    valid = ers == 0
    assert np.array_equal(rx[valid], erx[valid]) and valid.sum() > 100
    # depth from those matches (the oracle mutates its right list)
    right = [(int(s), (q[0], q[1])) for s, q in zip(ers, erx)]
    depths = m.get_depth_from_rectified_matches(lk, right)
    rs2, d = ctx.depth_from_rectified_matches(ls, lx[:, 0], ers, erx[:, 0])
    assert np.array_equal(rs2, np.array([s for s, _ in right], np.int32))
    assert np.array_equal(d, np.array(depths))


def test_median_disparity():
    p, rig, ctx = H.euroc_setup(batch=1, max_keypoints=2048)       # room for the > 1024-match (radix-select) path
    rng = np.random.default_rng(21)
    for n, m in ((50, 31), (400, 400), (1300, 1250), (10, 1)):       # > 1024 matches: the radix-select path
        ref = rng.uniform(0, 750, (n, 2)).astype(np.float32)
        cur = (ref + rng.normal(0, 3, (n, 2))).astype(np.float32)
        idx = rng.permutation(n)[:m]
        matches = np.stack([idx, rng.permutation(idx)], 1).astype(np.int32)
        ok, med = ofe.compute_median_disparity(ref, cur, [tuple(q) for q in matches])
        gok, gmed = ctx.compute_median_disparity(ref, cur, matches)
        assert gok == ok and gmed == med, (n, m, gmed, med)
    assert ctx.compute_median_disparity(ref, cur, np.zeros((0, 2), np.int32)) == (False, 0.0)
    ctx.close()


def test_point3_and_covariance(env):
    o, ctx = env["orig"], env["ctx"]
    rng = np.random.default_rng(4)
    n = 200
    uL = rng.uniform(50, 700, n).astype(np.float32)
    disp = rng.uniform(2, 60, n).astype(np.float32)
    v = rng.uniform(20, 460, n).astype(np.float32)
    left, right = np.stack([uL, v], 1), np.stack([uL - disp, v], 1)
    p3 = rng.uniform(-3, 3, (n, 3))
    calib = (o.fx, o.fy, o.cx, o.cy, o.baseline)
    Rm = cv2.Rodrigues(np.array([0.02, -0.05, 0.03]))[0]
    for R in (None, Rm):
        gp, gc = ctx.point3_and_covariance(left, right, p3, R)
        for i in range(n):
            ep, ec = ors.get_point3_and_covariance(left[i], right[i], p3[i], calib, R)
            assert np.array_equal(gp[i], ep)
            assert np.abs(gc[i] - ec).max() <= 1e-12 * max(1.0, np.abs(ec).max())


def test_detect_with_detection_mask(env):
    """Frame::detection_mask_ (FeatureDetector.cpp:186-189): a caller-supplied mask replaces the all-255 one; the
    circles around tracked keypoints are still drawn into it."""
    p, o, ctx = env["p"], env["orig"], env["ctx"]
    g, lefts, rights = H.golden()
    L = lefts[0]
    det = ofe.FeatureDetector(p)
    mask = np.full(L.shape, 255, np.uint8)
    mask[:, 300:520] = 0                   # a vertical band without detections
    mask[100:180, :] = 0
    fr = ofe.Frame(0, 0, L, o.left)
    fr.keypoints = [(np.float32(100.0), np.float32(200.0)), (np.float32(600.5), np.float32(400.25))]
    fr.landmarks = [5, -1]
    fr.landmarks_age, fr.scores = [1, 1], [0.0, 0.0]

    class MaskedDetector(ofe.FeatureDetector):
        def build_mask(self, frame):
            m = mask.copy()
            for kp, lmk in zip(frame.keypoints, frame.landmarks):
                if lmk != -1:
                    c = (int(np.rint(np.float32(kp[0]))), int(np.rint(np.float32(kp[1]))))
                    cv2.circle(m, c, self.p.min_distance, 0, cv2.FILLED)
            return m
    want = np.array(MaskedDetector(p).detect_corners(fr, 120), np.float32).reshape(-1, 2)
    got = ctx.detect_masked(L, mask, fr.keypoints, fr.landmarks, need=120)
    assert len(got) == len(want) and len(want) > 50
    assert np.abs(got - want).max() <= 1e-3
    assert not ((got[:, 0] > 300.5) & (got[:, 0] < 518.5)).any()


def test_forced_keyframe_sequence():
    """Frame::isKeyframe_ set by the caller (VisionImuFrontend.cpp:207-209): the forced frame becomes a keyframe in the
    stream that asked for it and only there."""
    from test_gpu_sequence import compare_packet, packet_ok
    N, B = 8, 2
    p, rig, ctx = H.euroc_setup(batch=B)
    orig = StereoRig(CameraParams.euroc_left(), CameraParams.euroc_right())
    s, fr = H.synth_frames(N, seed=20240)
    fes = [ofe.StereoFrontend(p, orig) for _ in range(B)]
    lkf, ok, kf_flags = [0] * B, True, [[], []]
    for k, f in enumerate(fr):
        force = [1 if (b == 0 and k in (2, 3)) else 0 for b in range(B)]
        if any(force):
            ctx.force_keyframe(force)
        Rs = [s.kf_rotation(lkf[b], k) for b in range(B)]
        pks = ctx.step([f.left] * B, [f.right] * B, [f.timestamp] * B, np.array(Rs))
        for b in range(B):
            sf = ofe.StereoFrame.make(k, f.timestamp, f.left, f.right, orig)
            sf.left_frame.is_keyframe = bool(force[b])
            o = fes[b].spin(sf, Rs[b])
            rec = compare_packet("forced/s%d/f%d" % (b, k), pks[b], o)
            rec["ok"] = packet_ok(rec)
            H.diag("sequence", **rec)
            ok &= rec["ok"]
            kf_flags[b].append(bool(pks[b]["is_keyframe"]))
            if o.is_keyframe:
                lkf[b] = k
    ctx.close()
    assert ok
    assert kf_flags[0][2] and kf_flags[0][3] and not (kf_flags[1][2] and kf_flags[1][3])
