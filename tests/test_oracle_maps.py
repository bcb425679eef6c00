"""CPU pin of the rectification-map formula the CUDA code evaluates (oracle/maps.py == common.cuh: rect_map_at) against
cv2.initUndistortRectifyMap on the reference's shipped rigs (tests/golden/rigs.json) and the RGB-D test camera: a coarse
grid everywhere plus the pixels where the fused operations decide the last bit."""
import cv2
import numpy as np
import pytest

import helpers as H
from kimera_vio_b200.params import CameraParams
from oracle import maps as om
from oracle.rig import StereoRig
from test_oracle_rgbd import rgbd_pair


def _check(cam, R, P, mx, my, pixels):
    bad = []
    for u, v in pixels:
        gx, gy = om.radtan_map_at(cam.K, cam.D, R, P, u, v)
        if gx.view(np.int32) != mx[v, u].view(np.int32) or gy.view(np.int32) != my[v, u].view(np.int32):
            bad.append((u, v, float(gx), float(mx[v, u]), float(gy), float(my[v, u])))
    return bad


def _pixels(W, Hh, extra_rows=(), step=37):
    px = [(u, v) for v in range(0, Hh, step) for u in range(0, W, step + 4)]
    px += [(0, v) for v in range(0, Hh, 3)] + [(u, 0) for u in range(0, W, 3)]          # zero crossings of a centred camera
    px += [(W - 1, v) for v in range(0, Hh, 29)] + [(u, Hh - 1) for u in range(0, W, 31)]
    for r in extra_rows:
        px += [(u, r) for u in range(0, W, 5)]
    return px


@pytest.mark.parametrize("rig_name", ["Euroc", "uHumans2", "uHumans1", "D455"])
def test_radtan_map_formula_matches_cv2(rig_name):
    p, left, right = H.shipped_rig(rig_name)
    o = StereoRig(left, right)
    px = _pixels(o.W, o.H, extra_rows=(15,))                    # row 15 of uHumans2: a f32 tie
    assert _check(left, o.R1, o.P1, o.map_lx, o.map_ly, px) == []
    bad = _check(right, o.R2, o.P2, o.map_rx, o.map_ry, px)
    # D455 right: one value of -1.16e-6 px (a zero crossing under distortion) differs in its last f32 bit, 1e-13 px
    assert len(bad) <= (1 if rig_name == "D455" else 0) and all(abs(b[4] - b[5]) < 1e-12 and abs(b[2] - b[3]) < 1e-12 for b in bad), bad


def test_radtan_map_formula_mono_camera_zero_crossings():
    """Camera (Camera.cpp:29-47): R = I, P = K with zero distortion: column 0 / row 0 of the map are residues of ~1e-14 px."""
    _, _, cam = rgbd_pair()
    mx, my = cv2.initUndistortRectifyMap(cam.K, cam.D, np.eye(3, dtype=np.float32), cam.K, (cam.width, cam.height), cv2.CV_32FC1)
    P = np.hstack([cam.K, np.zeros((3, 1))])
    assert _check(cam, np.eye(3), P, mx, my, _pixels(cam.width, cam.height)) == []
    assert mx[5, 0] != 0.0 and abs(mx[5, 0]) < 1e-12                                   # the residue the fusions reproduce
