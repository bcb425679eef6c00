"""The equidistant (cv::fisheye) distortion model of the rig -- StereoCamera.cpp:350-373 (cv::fisheye::stereoRectify),
UndistorterRectifier.cpp:49-56 (cv::fisheye::undistortPoints) and :260-268 (cv::fisheye::initUndistortRectifyMap) -- on
the reference's own params/RealSenseIR camera and front-end files (values from tests/golden/rigs.json): rectification
maps, rectified images, sparse undistortion in the three forms the path uses, the left-keypoint check / distort-unrectify
look-ups, and the whole front-end on a synthetic stream rendered through that fisheye rig, against the oracle (cv2)."""
import cv2
import numpy as np
import pytest

import helpers as H
from kimera_vio_b200 import lib as kl
from kimera_vio_b200.rig import MonoRigSetup, StereoRigSetup
from kimera_vio_b200.synth import SynthStream
from oracle import frontend as ofe
from oracle.mono import MonoFrontend
from oracle.rig import StereoRig
from test_gpu_sequence import run_sequence

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def env():
    p, left, right = H.shipped_rig("RealSenseIR")
    assert left.distortion_model == "equidistant" and right.distortion_model == "equidistant"
    rig = StereoRigSetup(left, right)
    cfg = kl.make_config(p, rig.W, rig.H, batch=1, sobel_cpu_tail_start=H.sobel_cpu_tail_start(rig.W))
    ctx = kl.Context(cfg, rig.to_c())
    s = SynthStream(left, right, rig.R1, seed=4242)
    yield dict(p=p, left=left, right=right, rig=rig, ctx=ctx, orig=StereoRig(left, right), synth=s)
    ctx.close()


def test_fisheye_maps_and_rectification(env):
    o, ctx = env["orig"], env["ctx"]
    for cam, (ex, ey) in enumerate(((o.map_lx, o.map_ly), (o.map_rx, o.map_ry))):
        mx, my = ctx.rectify_maps(cam)
        bad = int((mx.view(np.int32) != ex.view(np.int32)).sum() + (my.view(np.int32) != ey.view(np.int32)).sum())
        H.diag("fisheye_maps", cam=cam, mismatches=bad, max_abs=float(max(np.abs(mx - ex).max(), np.abs(my - ey).max())))
        assert bad == 0
    f = env["synth"].frame(0)
    gl, gr = ctx.rectify_pair(f.left, f.right)
    el, er = o.rectify_left(f.left), o.rectify_right(f.right)
    H.diag("fisheye_rectify", left_mismatches=int((gl != el).sum()), right_mismatches=int((gr != er).sum()))
    assert np.array_equal(gl, el) and np.array_equal(gr, er)


def test_fisheye_undistort_points_and_lookups(env):
    o, ctx = env["orig"], env["ctx"]
    W, Hh = env["rig"].W, env["rig"].H
    rng = np.random.default_rng(11)
    pts = np.stack([rng.uniform(0, W, 600), rng.uniform(0, Hh, 600)], 1).astype(np.float32)
    pts[0] = (np.float32(o.left.K[0, 2]), np.float32(o.left.K[1, 2]))     # the principal point: theta_d ~ 0
    for cam, camp, R, P in ((0, o.left, o.R1, o.P1), (1, o.right, o.R2, o.P2)):
        for useR, useP in ((False, False), (True, False), (True, True)):
            e = cv2.fisheye.undistortPoints(pts.reshape(-1, 1, 2), camp.K, camp.D, R=R if useR else None,
                                            P=P if useP else None).reshape(-1, 2)
            g = ctx.undistort_keypoints(cam, useR, useP, pts)
            bad = int((g != e).sum())
            H.diag("fisheye_undistort", cam=cam, useR=useR, useP=useP, mismatches=bad, max_abs=float(np.abs(g - e).max()))
            assert np.abs(g - e).max() <= 1e-3
            assert bad == 0
    v = ctx.bearing_vectors(pts)
    e = np.array(ofe.get_bearing_vectors([tuple(q) for q in pts], o.left, o.R1))
    assert np.abs(v - e).max() < 1e-12
    # UndistorterRectifier::undistortRectifyKeypoints + checkUndistortedRectifiedLeftKeypoints (the left-keypoint path)
    st, xy = ctx.undistort_rectify_left_keypoints(pts)
    exp = ofe.undistort_rectify_left_keypoints([tuple(q) for q in pts], o, 2.0)
    est = np.array([a for a, _ in exp], np.int32)
    exy = np.array([b for _, b in exp], np.float32).reshape(-1, 2)
    H.diag("fisheye_left_keypoints", n_valid=int((est == ofe.KP_VALID).sum()), status_mismatches=int((st != est).sum()),
           max_abs=float(np.abs(xy - exy).max()))
    assert np.array_equal(st, est) and np.array_equal(xy, exy)
    assert (est == ofe.KP_VALID).sum() > 300
    # distortUnrectifyKeypoints: the float maps at the rounded rectified position (UndistorterRectifier.cpp:213-228)
    for cam, (mx, my) in enumerate(((o.map_lx, o.map_ly), (o.map_rx, o.map_ry))):
        g = ctx.distort_unrectify_keypoints(cam, est, exy)
        r = np.clip(np.floor(exy.astype(np.float64) + 0.5).astype(np.int64), 0, [W - 1, Hh - 1])   # C round(), coordinates >= 0
        e = np.where((est == ofe.KP_VALID)[:, None], np.stack([mx[r[:, 1], r[:, 0]], my[r[:, 1], r[:, 0]]], 1), 0.0).astype(np.float32)
        assert np.array_equal(g, e)


def test_fisheye_sequence_realsense_ir(env):
    """params/RealSenseIR (640x480 equidistant, 5 cm baseline, maxFeatureAge 50, min_distance 8, keyframes every 0.1 s):
    12 frames of the whole stereo front-end, every packet equal to the oracle's."""
    p, left, right, rig = env["p"], env["left"], env["right"], env["rig"]
    cfg = kl.make_config(p, rig.W, rig.H, batch=1, sobel_cpu_tail_start=H.sobel_cpu_tail_start(rig.W))
    ctx = kl.Context(cfg, rig.to_c())
    s = env["synth"]
    fr = [s.frame(k) for k in range(12)]
    fe = ofe.StereoFrontend(p, StereoRig(left, right))
    ok = run_sequence(ctx, [fe], [[(f.left, f.right, f.timestamp) for f in fr]],
                      lambda b, k, l: s.kf_rotation(l, k), "rig_RealSenseIR")
    ctx.close()
    assert ok


def test_fisheye_mono_sequence(env):
    """The mono front-end (Camera.cpp:29-47: P = K, R = I) on the RealSenseIR left camera: keypoints_undistorted_ come
    from cv::fisheye::undistortPoints and the fisheye maps."""
    p, cam = env["p"], env["left"]
    rig = MonoRigSetup(cam)
    cfg = kl.make_config(p, rig.W, rig.H, batch=1, sobel_cpu_tail_start=H.sobel_cpu_tail_start(rig.W), mono=True)
    ctx = kl.Context(cfg, rig.to_c())
    s = env["synth"]
    fe = MonoFrontend(p, cam)
    lkf, bad, n_kf = 0, [], 0
    for k in range(8):
        f = s.frame(k)
        R = s.kf_rotation(lkf, k)
        pk = ctx.step([f.left], [f.left], [f.timestamp], np.array([R]))[0]
        o, is_kf, smart = fe.spin(ofe.Frame(k, f.timestamp, f.left, cam), R)
        ok = pk["n"] == len(o.keypoints) and bool(pk["is_keyframe"]) == bool(is_kf)
        rec = dict(k=k, n=(int(pk["n"]), len(o.keypoints)), kf=(int(pk["is_keyframe"]), int(is_kf)))
        if ok and pk["n"]:
            kp = np.array(o.keypoints, np.float32).reshape(-1, 2)
            rec["kp_err"] = float(np.abs(np.stack([pk["kp_x"], pk["kp_y"]], 1) - kp).max())
            ok &= rec["kp_err"] <= 1e-3
            ok &= np.array_equal(pk["landmark"], np.array(o.landmarks, np.int64))
            rec["versor_err"] = float(np.abs(pk["versor"] - np.array(o.versors).reshape(-1, 3)).max())
            ok &= rec["versor_err"] < 1e-5
        ok &= (pk["mono_status"] == fe.mono_status) or k == 0
        if ok and is_kf:
            n_kf += 1
            us = np.array([st for st, _ in o.keypoints_undistorted], np.int32)
            ux = np.array([q for _, q in o.keypoints_undistorted], np.float32).reshape(-1, 2)
            ok &= np.array_equal(pk["left_status"], us)
            rec["undist_err"] = float(np.abs(np.stack([pk["left_rect_x"], pk["left_rect_y"]], 1) - ux).max())
            ok &= rec["undist_err"] <= 2e-3
        rec["ok"] = bool(ok)
        H.diag("fisheye_mono_sequence", **rec)
        if not ok:
            bad.append(rec)
        if is_kf:
            lkf = k
    ctx.close()
    assert not bad, bad[:3]
    assert n_kf >= 3


def test_unknown_distortion_model_rejected(env):
    """The omni model (Camera::UndistortKeypointsOmni) is not built: kvfe_create refuses it instead of ignoring it."""
    r = env["rig"].to_c()
    r.distortion_model = 2
    cfg = kl.make_config(env["p"], env["rig"].W, env["rig"].H, batch=1)
    with pytest.raises(kl.KvfeError):
        kl.Context(cfg, r)
