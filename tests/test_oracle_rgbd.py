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
