"""Row f2, RGB-D at the stage level: DepthFrame::getDetectionMask (DepthFrame.cpp:75-98) and RgbdFrame::fillStereoFrame
(RgbdFrame.cpp:52-115) through the C-ABI against oracle/rgbd.py (itself pinned on the reference's testDepthFrame /
testRgbdFrame known answers) on the reference's real RGB-D pair (tests/golden/rgbd_pair.npz: depth_img_0.tiff CV_32FC1 +
left_img_0.png, camera sensorLeft.yaml), on a millimetre CV_16UC1 copy of it, with NaN / inf / zero depths injected, and
on adversarial keypoints (truncation at the image border, uR < 0).  Everything is compared exactly."""
import cv2
import numpy as np
import pytest

import helpers as H
from kimera_vio_b200 import lib as kl
from kimera_vio_b200.params import FrontendParams
from kimera_vio_b200.rig import RgbdRigSetup
from oracle import frontend as ofe
from oracle import rgbd as org
from oracle.mono import MonoCamera
from test_oracle_rgbd import rgbd_pair

pytestmark = pytest.mark.gpu
f32 = np.float32


@pytest.fixture(scope="module")
def env():
    depth, left, cam = rgbd_pair()
    rig = RgbdRigSetup(cam)
    p = FrontendParams.euroc()
    cfg = kl.make_config(p, rig.W, rig.H, batch=1, sobel_cpu_tail_start=H.sobel_cpu_tail_start(rig.W), mono=True)
    ctx = kl.Context(cfg, rig.to_c())
    yield dict(depth=depth, left=left, cam=cam, rig=rig, p=p, ctx=ctx, mono=MonoCamera(cam))
    ctx.close()


def _dp(cam, dtype, **over):
    d = {k: v for k, v in cam.depth.items() if k != "is_registered"}
    d.update(over)
    return d, kl.make_depth_params(dtype, **d)


def _spoiled(depth):
    d = depth.copy()
    d[10:40, 100:200] = np.nan
    d[50:60, 300:420] = np.inf
    d[200:230, 500:640] = 0.0
    d[300:310, 10:90] = -1.0
    return d


def test_depth_detection_mask(env):
    ctx, cam = env["ctx"], env["cam"]
    inf = float("inf")
    d32 = _spoiled(env["depth"])
    cases = [dict(min_depth=0.0, max_depth=inf), dict(min_depth=inf, max_depth=inf), dict(min_depth=0.2, max_depth=3.0),
             dict(), dict(min_depth=0.5, max_depth=6.0, depth_to_meters=0.5)]
    for over in cases:
        d, dp = _dp(cam, np.float32, **over)
        got, want = ctx.depth_detection_mask(d32, dp), org.get_detection_mask(d32, d)
        H.diag("depth_mask", dtype="f32", **{k: str(v) for k, v in over.items()}, mean=float(want.mean()), mismatches=int((got != want).sum()))
        assert np.array_equal(got, want)
    mm = np.clip(np.nan_to_num(d32, nan=0.0, posinf=70000.0, neginf=0.0) * 1000.0, 0, 65535).astype(np.uint16)
    for over in (dict(min_depth=0.2, max_depth=3.0), dict(), dict(min_depth=0.0, max_depth=100.0), dict(min_depth=3.0007, max_depth=3.9993)):
        d, dp = _dp(cam, np.uint16, depth_to_meters=0.001, **over)
        got, want = ctx.depth_detection_mask(mm, dp), org.get_detection_mask(mm, d)
        H.diag("depth_mask", dtype="u16", **{k: str(v) for k, v in over.items()}, mean=float(want.mean()), mismatches=int((got != want).sum()))
        assert np.array_equal(got, want)
        assert 0 < want.mean() <= 255
    # the detection the RGB-D front-end runs with that mask (RgbdVisionImuFrontend.cpp:196-199): kvfe_detect_masked
    d, dp = _dp(cam, np.float32, min_depth=0.2, max_depth=6.0)
    mask = ctx.depth_detection_mask(d32, dp)

    class DepthMaskedDetector(ofe.FeatureDetector):
        def build_mask(self, frame):
            return mask.copy()
    fr = ofe.Frame(0, 0, env["left"], cam)
    want = np.array(DepthMaskedDetector(env["p"]).detect_corners(fr, 200), np.float32).reshape(-1, 2)
    got = ctx.detect_masked(env["left"], mask, [], [], need=200)
    assert len(got) == len(want) and len(want) > 50 and np.abs(got - want).max() <= 1e-3


def _compare_fill(tag, ctx, depth, d, dp, cam, mono, kps, left, versors):
    er, ed, ep, ek = org.fill_stereo_frame(depth, _with_depth(cam, d), kps, left, versors, mono.map_lx, mono.map_ly)
    ls = np.array([s for s, _ in left], np.int32)
    lxy = np.array([q for _, q in left], np.float32).reshape(-1, 2)
    rs, rxy, dep, p3, rk = ctx.rgbd_fill_stereo_frame(depth, dp, np.array(kps, np.float32).reshape(-1, 2), ls, lxy, np.array(versors))
    ers = np.array([s for s, _ in er], np.int32)
    erxy = np.array([q for _, q in er], np.float32).reshape(-1, 2)
    rec = dict(tag=tag, n=len(kps), n_valid=int((ers == ofe.KP_VALID).sum()), n_no_depth=int((ers == ofe.KP_NO_DEPTH).sum()),
               status_mismatches=int((rs != ers).sum()), right_mismatches=int((rxy != erxy).sum()),
               depth_mismatches=int((dep != np.array(ed)).sum()), p3d_mismatches=int((p3 != np.array(ep).reshape(-1, 3)).sum()),
               right_kp_mismatches=int((rk != np.array(ek, np.float32).reshape(-1, 2)).sum()),
               right_kp_max_err=float(np.abs(rk - np.array(ek, np.float32).reshape(-1, 2)).max()))
    H.diag("rgbd_fill", **rec)
    assert rec["status_mismatches"] == 0 and rec["right_mismatches"] == 0 and rec["depth_mismatches"] == 0
    assert rec["p3d_mismatches"] == 0
    # right_frame_.keypoints_ are raw map values, incl. column 0 / row 0 of this zero-distortion camera where the map is a
    # residue of 1e-14 px that only OpenCV's fused operation order reproduces (common.cuh: rect_map_at, tests/test_oracle_maps.py)
    assert rec["right_kp_mismatches"] == 0
    return rec


def _with_depth(cam, d):
    import dataclasses
    return dataclasses.replace(cam, depth=dict(d, is_registered=True))


def test_rgbd_fill_stereo_frame_real_pair(env):
    ctx, cam, mono, left_img = env["ctx"], env["cam"], env["mono"], env["left"]
    c = cv2.goodFeaturesToTrack(left_img, 400, 0.001, 10).reshape(-1, 2).astype(np.float32)
    kps = [(f32(x), f32(y)) for x, y in c]
    left = ofe.undistort_rectify_left_keypoints(kps, mono, 2.0)          # Camera::undistortKeypoints
    versors = ofe.get_bearing_vectors(kps, cam, None)
    assert sum(1 for s, _ in left if s == ofe.KP_VALID) > 200
    d32 = _spoiled(env["depth"])
    seen_no_depth = 0
    for tag, over in (("default", dict()), ("min2.5", dict(min_depth=2.5)), ("wide_baseline", dict(virtual_baseline=5.0)),
                      ("scaled", dict(depth_to_meters=0.5, min_depth=1.0))):
        d, dp = _dp(cam, np.float32, **over)
        rec = _compare_fill("f32/" + tag, ctx, d32, d, dp, cam, mono, kps, left, versors)
        seen_no_depth += rec["n_no_depth"]
        assert rec["n_valid"] > 5
    assert seen_no_depth > 50
    mm = np.clip(np.nan_to_num(d32, nan=0.0, posinf=70000.0, neginf=0.0) * 1000.0, 0, 65535).astype(np.uint16)
    d, dp = _dp(cam, np.uint16, depth_to_meters=0.001, min_depth=0.3)
    rec = _compare_fill("u16/mm", ctx, mm, d, dp, cam, mono, kps, left, versors)
    assert rec["n_valid"] > 100 and rec["n_no_depth"] > 0


def test_rgbd_fill_stereo_frame_adversarial(env):
    """Keypoints on and beyond the image border (static_cast<int> truncation: -0.5 reads column 0, W - 0.5 reads the last
    column, W is outside), every left status, arbitrary versors, rectified x smaller than the disparity (uR < 0)."""
    ctx, cam, mono = env["ctx"], env["cam"], env["mono"]
    W, Hh = env["rig"].W, env["rig"].H
    rng = np.random.default_rng(5)
    n = 700
    kx = rng.uniform(-1.5, W + 1.5, n).astype(np.float32)
    ky = rng.uniform(-1.5, Hh + 1.5, n).astype(np.float32)
    kx[:6] = (-0.5, -1.0, W - 0.5, W, 0.0, W - 1)
    ky[:6] = (0.0, 5.0, Hh - 0.5, 7.0, Hh, Hh - 1)
    kps = [(x, y) for x, y in zip(kx, ky)]
    st = rng.choice([0, 0, 0, 0, 1, 2, 3, 4], n).astype(np.int32)
    st[:6] = 0
    lx = rng.uniform(0, W - 1, n).astype(np.float32)
    lx[::7] = rng.uniform(0, 3, len(lx[::7])).astype(np.float32)        # small x: uR < 0 for near points
    ly = rng.uniform(0, Hh - 1, n).astype(np.float32)
    left = [(int(s), (x, y)) for s, x, y in zip(st, lx, ly)]
    v = rng.normal(0, 0.4, (n, 3))
    v[:, 2] = rng.uniform(0.5, 1.0, n)
    v /= np.linalg.norm(v, axis=1)[:, None]
    versors = [q for q in v]
    d32 = _spoiled(env["depth"])
    d, dp = _dp(cam, np.float32, virtual_baseline=0.05, min_depth=0.0)
    rec = _compare_fill("adversarial/f32", ctx, d32, d, dp, cam, mono, kps, left, versors)
    assert rec["n_valid"] > 100 and rec["n_no_depth"] > 30
    # no keypoints at all: nothing is written, no error (testRgbdFrame.cpp:108-117)
    out = ctx.rgbd_fill_stereo_frame(d32, dp, np.zeros((0, 2), np.float32), np.zeros(0, np.int32), np.zeros((0, 2), np.float32), np.zeros((0, 3)))
    assert all(len(a) == 0 for a in out)
    with pytest.raises(kl.KvfeError):
        ctx.depth_detection_mask(d32, kl.DepthParams(7, 0.1, 1.0, 0.0, 10.0))
    with pytest.raises(kl.KvfeError):
        ctx.depth_detection_mask(d32, kl.DepthParams(1, 0.0, 1.0, 0.0, 10.0))
