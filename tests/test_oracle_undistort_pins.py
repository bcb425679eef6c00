"""CPU pins of the oracle's UndistorterRectifier on the reference's own property tests, with the reference's test camera
(tests/data/ForStereoFrame/sensor{Left,Right}.yaml, carried by tests/golden/tracker_scenes.npz):
  * tests/testUndistortRectifier.cpp:87-145  undistortRectifyKeypoints -> checkUndistortedRectifiedLeftKeypoints ->
    distortUnrectifyKeypoints returns to the original pixel within 1 px for every VALID point of an 8 x 10 grid;
  * tests/testUndistortRectifier.cpp:152-220 distortUnrectifyKeypoints equals K * distort(R1^T P1^-1 pt) (gtsam::Cal3DS2::
    uncalibrate) within 1e-3 px on the grid, (0, 0) for non-VALID points.
The same two properties are asserted for the equidistant model on the reference's params/RealSenseIR rig (its own fisheye
tests, testUndistortRectifier.cpp:225-330, are DISABLED upstream: "ref image seems incorrect")."""
import cv2
import numpy as np

import helpers as H
from oracle import frontend as ofe
from oracle.rig import StereoRig
from test_oracle_ransac import seeded_scenes


def _grid(W, Hh, rows=8, cols=10):
    return [(np.float32(W // (cols - 1) * c), np.float32(Hh // (rows - 1) * r)) for r in range(rows) for c in range(cols)]


def _round_trip(o):
    gt = _grid(o.W, o.H)
    rect = ofe.undistort_rectify_left_keypoints(gt, o)        # undistortRectifyKeypoints + check (tol 2.0, the default)
    n_valid = 0
    for (st, (x, y)), (gx, gy) in zip(rect, gt):
        if st != ofe.KP_VALID:
            continue
        n_valid += 1
        back = (o.map_lx[ofe.c_round(y), ofe.c_round(x)], o.map_ly[ofe.c_round(y), ofe.c_round(x)])   # distortUnrectifyKeypoints
        assert abs(back[0] - gx) <= 1 and abs(back[1] - gy) <= 1, ((gx, gy), (x, y), back)
    return n_valid, len(gt)


def test_undistort_rectify_keypoints_round_trip_reference_camera():
    _, left, right = seeded_scenes()
    n_valid, n = _round_trip(StereoRig(left, right))
    assert n == 80 and n_valid >= 40


def test_undistort_rectify_keypoints_round_trip_equidistant():
    _, left, right = H.shipped_rig("RealSenseIR")
    n_valid, n = _round_trip(StereoRig(left, right))
    assert n == 80 and n_valid >= 40


def _cal3ds2_uncalibrate(cam, x, y):
    """gtsam::Cal3DS2::uncalibrate (radial-tangential, skew 0)."""
    k1, k2, p1, p2 = cam.distortion[:4]
    fx, fy, u0, v0 = cam.intrinsics
    xy, xx, yy = x * y, x * x, y * y
    rr = xx + yy
    g = 1.0 + k1 * rr + k2 * rr * rr
    dx = 2.0 * p1 * xy + p2 * (rr + 2.0 * xx)
    dy = 2.0 * p2 * xy + p1 * (rr + 2.0 * yy)
    return fx * (g * x + dx) + u0, fy * (g * y + dy) + v0


def test_distort_unrectify_keypoints_reference_camera():
    _, left, right = seeded_scenes()
    o = StereoRig(left, right)
    pts = _grid(o.W, o.H)
    rect = [((ofe.KP_NO_RIGHT_RECT if (i // 10 + i % 10) % 2 == 0 else ofe.KP_VALID), p) for i, p in enumerate(pts)]
    # the left camera's maps through the generic look-up
    got = []
    for st, (x, y) in rect:
        if st == ofe.KP_VALID:
            xx, yy = min(ofe.c_round(x), o.W - 1), min(ofe.c_round(y), o.H - 1)      # the grid's last row / column is W, H
            if ofe.c_round(x) >= o.W or ofe.c_round(y) >= o.H:
                got.append(None)                                                       # outside the map (the reference reads past it)
                continue
            got.append((float(o.map_lx[yy, xx]), float(o.map_ly[yy, xx])))
        else:
            got.append((0.0, 0.0))
    P1_inv = np.linalg.inv(o.P1[:3, :3])
    n_checked = 0
    for (st, (x, y)), g in zip(rect, got):
        if st != ofe.KP_VALID:
            assert g == (0.0, 0.0)
            continue
        if g is None:
            continue
        xn = o.R1.T @ P1_inv @ np.array([float(x), float(y), 1.0])
        ex, ey = _cal3ds2_uncalibrate(left, xn[0] / xn[2], xn[1] / xn[2])
        assert abs(g[0] - ex) < 1e-3 and abs(g[1] - ey) < 1e-3, ((x, y), g, (ex, ey))
        n_checked += 1
    assert n_checked >= 30


def test_get_bearing_vector_grid_reference():
    """tests/testFrame.cpp:101-139 (sensor.yaml == the Euroc left camera): GetBearingVector has unit norm and, distorted
    again with Cal3DS2::uncalibrate, lands within 0.5 px of the pixel it came from, on an 8 x 8 grid over 752 x 480."""
    from kimera_vio_b200.params import CameraParams
    cam = CameraParams.euroc_left()
    pts = [(np.float32(c * 752 // 7), np.float32(r * 480 // 7)) for r in range(8) for c in range(8)]
    for (px, py), v in zip(pts, ofe.get_bearing_vectors(pts, cam, None)):
        assert abs(np.linalg.norm(v) - 1.0) < 4e-16
        ex, ey = _cal3ds2_uncalibrate(cam, v[0] / v[2], v[1] / v[2])
        assert np.hypot(ex - float(px), ey - float(py)) < 0.5
