"""CPU pins of the equidistant rig model (no GPU): the scalar restatement the CUDA code follows (oracle/fisheye.py)
against cv2.fisheye bit for bit on the reference's params/RealSenseIR rig (tests/golden/rigs.json), the rig marshalling,
and the synthetic renderer's fisheye camera against cv2."""
import cv2
import numpy as np

import helpers as H
from kimera_vio_b200 import lib as kl
from kimera_vio_b200.rig import MonoRigSetup, StereoRigSetup
from kimera_vio_b200.synth import SynthStream, _undistort_grid
from oracle import fisheye as fe
from oracle.rig import StereoRig


def _rig():
    p, left, right = H.shipped_rig("RealSenseIR")
    return p, left, right, StereoRig(left, right)


def test_fisheye_maps_restatement_bit_exact():
    _, left, right, o = _rig()
    for cam, R, P, (ex, ey) in ((left, o.R1, o.P1, (o.map_lx, o.map_ly)), (right, o.R2, o.P2, (o.map_rx, o.map_ry))):
        mx, my = fe.init_undistort_rectify_map(cam.K, cam.D, R, P, (o.W, o.H))
        assert np.array_equal(mx.view(np.int32), ex.view(np.int32))
        assert np.array_equal(my.view(np.int32), ey.view(np.int32))
    # the mono camera's maps: R = I, P = K (Camera.cpp:29-47)
    ex, ey = cv2.fisheye.initUndistortRectifyMap(left.K, left.D, np.eye(3), left.K, (o.W, o.H), cv2.CV_32FC1)
    mx, my = fe.init_undistort_rectify_map(left.K, left.D, np.eye(3), np.hstack([left.K, np.zeros((3, 1))]), (o.W, o.H))
    assert np.array_equal(mx.view(np.int32), ex.view(np.int32)) and np.array_equal(my.view(np.int32), ey.view(np.int32))


def test_fisheye_undistort_points_restatement_bit_exact():
    _, left, right, o = _rig()
    rng = np.random.default_rng(0)
    near = np.stack([rng.uniform(-50, o.W + 50, 1500), rng.uniform(-50, o.H + 50, 1500)], 1).astype(np.float32)
    near[0] = (np.float32(left.K[0, 2]), np.float32(left.K[1, 2]))
    far = np.stack([rng.uniform(-3000, 3000, 1500), rng.uniform(-3000, 3000, 1500)], 1).astype(np.float32)
    n_failed = 0
    for D in (left.D, np.array([[-0.3, 0.2, -0.5, 0.1]]), np.array([[0.9, -2.0, 3.0, -1.0]])):
        for pts in (near, far):
            for R, P in ((o.R1, o.P1), (o.R1, None), (None, None)):
                e = cv2.fisheye.undistortPoints(pts.reshape(-1, 1, 2), left.K, D, R=R, P=P).reshape(-1, 2)
                g = fe.undistort_points(pts, left.K, D, R, P)
                assert np.array_equal(g.view(np.int32), e.view(np.int32))
                n_failed += int((e[:, 0] == -1000000.0).sum())
    assert n_failed > 100          # the non-convergent / sign-flip branch is exercised


def test_equidistant_rig_marshalling_and_rectification_geometry():
    p, left, right, o = _rig()
    rig = StereoRigSetup(left, right)
    c = rig.to_c()
    assert c.distortion_model == kl.DISTORTION_MODELS["equidistant"] == 1 and c.reserved == 0
    assert list(c.D_left) == list(left.distortion[:4])
    assert np.array_equal(np.array(c.P1).reshape(3, 4), o.P1) and abs(c.baseline - o.baseline) == 0
    assert 0.045 < rig.baseline < 0.055                              # params/RealSenseIR: 5 cm
    m = MonoRigSetup(left).to_c()
    assert m.distortion_model == 1 and np.array_equal(np.array(m.P1).reshape(3, 4)[:, :3], left.K)
    assert kl.make_rig(*_euroc_pair(), np.eye(3), np.eye(3), np.eye(3, 4), np.eye(3, 4), 0.11).distortion_model == 0


def _euroc_pair():
    from kimera_vio_b200.params import CameraParams
    return CameraParams.euroc_left(), CameraParams.euroc_right()


def test_synth_fisheye_camera_matches_cv2():
    _, left, right, o = _rig()
    g = _undistort_grid(left)
    ys, xs = np.mgrid[0:o.H:37, 0:o.W:41]
    pts = np.stack([xs.ravel(), ys.ravel()], 1).astype(np.float32)
    e = cv2.fisheye.undistortPoints(pts.reshape(-1, 1, 2), left.K, left.D).reshape(-1, 2)
    assert np.abs(g[ys.ravel(), xs.ravel()] - e).max() < 1e-5
    s = SynthStream(left, right, o.R1, seed=4242)
    f = s.frame(0)
    assert f.left.shape == (o.H, o.W) and f.left.std() > 10 and f.right.std() > 10
