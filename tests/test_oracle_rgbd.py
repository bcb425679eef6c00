"""CPU pins of the RGB-D oracle (oracle/rgbd.py) on the reference's own known answers:
tests/testDepthFrame.cpp:129-173 (GetDepthAtPoint, CV_32FC1 and CV_16UC1), :58-96 (DetectionMask on
tests/data/ForRgbd/depth_img_0.tiff, here tests/golden/rgbd_pair.npz) and tests/testRgbdFrame.cpp:84-172 (FillStereoFrame),
plus the host-side marshalling (kvfe_depth_params, the RGB-D rig)."""
import json
import math
import os

import cv2
import numpy as np
import pytest

import helpers as H
from kimera_vio_b200 import lib as kl
from kimera_vio_b200.params import CameraParams
from kimera_vio_b200.rig import RgbdRigSetup
from oracle import frontend as ofe
from oracle import rgbd as org

f32 = np.float32


def rgbd_pair():
    g = np.load(os.path.join(H.ROOT, "tests", "golden", "rgbd_pair.npz"))
    c = json.loads(str(g["camera"]))
    c["T_BS"] = np.asarray(c["T_BS"], np.float64)
    return g["depth"], g["left"], CameraParams(**c)


@pytest.mark.parametrize("dtype", [np.float32, np.uint16])
def test_get_depth_at_point_reference_cases(dtype):
    img = np.array([[1, 2], [3, 4]], dtype)                       # makeTestDepthFloat / makeTestDepthUINT16
    dp = {"depth_to_meters": 1.0, "min_depth": 0.1}
    for pt in ((-1.0, 0.0), (2.0, 0.0), (0.0, -1.0), (0.0, 2.0)):  # first case: invalid pixels
        assert math.isnan(org.get_depth_at_point(img, dp, pt))
    assert abs(org.get_depth_at_point(img, dp, (0.0, 0.0)) - 1.0) < 1e-9
    assert abs(org.get_depth_at_point(img, dict(dp, depth_to_meters=5.0), (0.0, 0.0)) - 5.0) < 1e-9
    assert math.isnan(org.get_depth_at_point(img, dict(dp, depth_to_meters=0.01), (0.0, 0.0)))
    # truncation, not rounding: (1.9, 0.9) reads pixel (1, 0); (-0.5, 0) truncates to column 0
    assert org.get_depth_at_point(img, dp, (1.9, 0.9)) == 2.0 and org.get_depth_at_point(img, dp, (-0.5, 0.0)) == 1.0


def test_detection_mask_reference_cases():
    depth, _, _ = rgbd_pair()
    base = {"depth_to_meters": 1.0}
    m = org.get_detection_mask(depth, dict(base, min_depth=0.0, max_depth=float("inf")))
    assert m.dtype == np.uint8 and m.mean() == 255.0                                      # everything allowed
    assert org.get_detection_mask(depth, dict(base, min_depth=float("inf"), max_depth=float("inf"))).mean() == 0.0
    avg = org.get_detection_mask(depth, dict(base, min_depth=0.2, max_depth=3.0)).mean()
    assert 0.0 < avg < 255.0
    # uint16 millimetres
    mm = np.clip(depth * 1000.0, 0, 65535).astype(np.uint16)
    m16 = org.get_detection_mask(mm, {"depth_to_meters": 0.001, "min_depth": 0.2, "max_depth": 3.0})
    assert abs(m16.mean() - avg) < 1.0


def test_fill_stereo_frame_reference_case():
    _, _, cam = rgbd_pair()
    cam.depth.update(depth_to_meters=1.0, min_depth=1.5, is_registered=True, virtual_baseline=float(f32(1.0e-1)))
    fx = f32(cam.intrinsics[0])
    depth = np.array([[1.0, fx * f32(2.0)], [3.0, fx * f32(4.0)]], np.float32)
    mx, my = cv2.initUndistortRectifyMap(cam.K, cam.D, np.eye(3, dtype=np.float32), cam.K, (cam.width, cam.height), cv2.CV_32FC1)
    # case 1: no features
    r, d, p, k = org.fill_stereo_frame(depth, cam, [], [], [], mx, my)
    assert r == [] and d == [] and p == [] and k == []
    # case 2: some features
    kps = [(0.0, 0.0), (1.0, 0.0), (0.0, 1.0), (1.0, 1.0)]
    FAILED_ARUN = 4
    left = [(ofe.KP_VALID, kps[0]), (ofe.KP_VALID, kps[1]), (FAILED_ARUN, kps[2]), (ofe.KP_VALID, kps[3])]
    r, d, p, k = org.fill_stereo_frame(depth, cam, kps, left, [np.ones(3)] * 4, mx, my)
    assert len(r) == len(d) == len(p) == len(k) == 4
    for got, exp in zip(d, (0.0, float(fx * f32(2.0)), 0.0, float(fx * f32(4.0)))):
        assert abs(got - exp) < 1e-9
    assert [s for s, _ in r] == [ofe.KP_NO_DEPTH, ofe.KP_VALID, FAILED_ARUN, ofe.KP_VALID]
    assert abs(r[1][1][0] - 0.95) < 1e-6 and abs(r[3][1][0] - 0.975) < 1e-6           # 1 - 0.1 / 2, 1 - 0.1 / 4
    assert np.allclose(p[1], float(fx * f32(2.0))) and np.allclose(p[0], 0.0)


def test_depth_params_and_rgbd_rig_marshalling():
    _, _, cam = rgbd_pair()
    assert abs(cam.depth["virtual_baseline"] - float(f32(0.3))) == 0 and cam.depth["max_depth"] == 10.0
    dp = kl.make_depth_params(np.float32, **{k: v for k, v in cam.depth.items() if k != "is_registered"})
    assert dp.depth_type == 1 and dp.virtual_baseline == f32(0.3) and dp.min_depth == 0.0
    assert kl.make_depth_params(np.uint16).depth_type == 0 and kl.make_depth_params(np.uint16).virtual_baseline == f32(1.0e-2)
    with pytest.raises(ValueError):
        kl.make_depth_params(np.float64)
    rig = RgbdRigSetup(cam)
    c = rig.to_c()
    assert c.baseline == float(f32(0.3)) and np.array_equal(np.array(c.P1).reshape(3, 4)[:, :3], cam.K) and c.distortion_model == 0
    with pytest.raises(ValueError):
        RgbdRigSetup(CameraParams.euroc_left())


def test_rgbd_frontend_oracle_on_synthetic_sequence():
    """oracle/rgbd.py: RgbdFrontend (RgbdVisionImuFrontend.cpp:183-395) over a synthetic RGB-D stream (the stereo test scene
    seen by one camera, with its metric depth image): the properties the reference's flow guarantees -- first frame is a
    keyframe with detections only where the depth mask allows, keyframes follow the time rule, the hallucinated right
    keypoints satisfy uR = uL - fx b / depth with the depth of the truncated raw pixel, 3-D points lie on the rendered
    planes, both RANSAC stages accept the rigid scene, smart measurements carry uR (fillSmartStereoMeasurements)."""
    import dataclasses
    from kimera_vio_b200.params import FrontendParams
    from kimera_vio_b200.rig import StereoRigSetup
    from kimera_vio_b200.synth import SynthStream
    p = FrontendParams.euroc()
    left, right = CameraParams.euroc_left(), CameraParams.euroc_right()
    cam = dataclasses.replace(left, depth={"virtual_baseline": float(f32(0.1)), "depth_to_meters": 1.0, "min_depth": 0.3,
                                           "max_depth": 4.0, "is_registered": True})
    s = SynthStream(left, right, np.eye(3), seed=99)
    fe = org.RgbdFrontend(p, cam)
    lkf, n_kf, statuses = 0, 0, []
    for k in range(13):
        f, depth = s.frame_with_depth(k)
        assert depth.shape == f.left.shape and 1.0 < depth[depth > 0].min() and depth.max() < 7.0
        R = s.kf_rotation(lkf, k)
        sf, is_kf, smart = fe.spin(k, f.timestamp, f.left, depth, R)
        lf = sf.left_frame
        assert len(lf.keypoints) == len(lf.landmarks) == len(lf.versors) == len(sf.left_keypoints_rectified)
        if is_kf:
            n_kf += 1
            lkf = k
            assert len(sf.right_keypoints_rectified) == len(lf.keypoints) == len(sf.keypoints_3d)
            fx_b = cam.intrinsics[0] * float(f32(0.1))
            n_valid = 0
            for i, (st, (rx, ry)) in enumerate(sf.right_keypoints_rectified):
                lst, (lx, ly) = sf.left_keypoints_rectified[i]
                if st != ofe.KP_VALID:
                    continue
                n_valid += 1
                d = float(depth[int(lf.keypoints[i][1]), int(lf.keypoints[i][0])])
                assert lst == ofe.KP_VALID and d >= 0.3 and sf.keypoints_depth[i] == d
                assert abs(float(rx) - (float(lx) - fx_b / d)) < 1e-4 and ry == ly
                assert abs(sf.keypoints_3d[i][2] - d) < 1e-9                       # versor * depth / versor.z
            assert n_valid > 150
            # detections respect the depth mask (max_depth 4 m cuts the far plane at 6 m) up to the sub-pixel refinement
            new = [kp for kp, age in zip(lf.keypoints, lf.landmarks_age) if age == 1]
            far = sum(1 for (x, y) in new if depth[int(y), int(x)] > 4.5)
            assert far <= len(new) * 0.05
            if k > 0:
                statuses.append((fe.mono_status, fe.stereo_status))
                assert len(smart) == sum(1 for l in lf.landmarks if l != -1)
                with_uR = sum(1 for m in smart if not math.isnan(m[2]))
                assert with_uR > 100
        else:
            assert smart == []
    assert n_kf >= 4
    assert all(m == ofe.VALID and st == ofe.VALID for m, st in statuses), statuses
