"""Edge-case sequences through the device-resident FSM vs the oracle: a textureless stereo pair in the
middle of a sequence (tracking into it, detection on it, every track lost on the next frame ->
re-detection without a keyframe, StereoVisionImuFrontend.cpp:312-323) and a sequence in which the
max_feature_age limit removes tracks (Tracker.cpp:150-160).
A sensor-noise-only pair is deliberately NOT a parity case for the stereo matcher: cv2 evaluates
TM_SQDIFF through a float DFT, so among hundreds of nearly tied shifts its arg-min differs from the
exact-integer one (measured: 2 of 286 keypoints on a sigma-2 noise pair); see DESIGN.md section 4.
(A first pair without any corner is not a parity case: the reference CHECK-fails in
StereoMatcher::sparseStereoReconstruction, "Call feature detection on left frame first".)"""
import numpy as np
import pytest

import helpers as H
from kimera_vio_b200.params import CameraParams
from oracle import frontend as ofe
from oracle.rig import StereoRig
from test_gpu_sequence import run_sequence

pytestmark = pytest.mark.gpu


def test_sequence_blank_frame():
    p, rig, ctx = H.euroc_setup(batch=1)
    orig = StereoRig(CameraParams.euroc_left(), CameraParams.euroc_right())
    s, fr = H.synth_frames(12, seed=9090)
    frames = [(f.left.copy(), f.right.copy(), f.timestamp) for f in fr]
    blank = np.full_like(frames[0][0], 128)
    frames[4] = (blank, blank.copy(), frames[4][2])          # textureless pair
    fe = ofe.StereoFrontend(p, orig)
    ok = run_sequence(ctx, [fe], [frames], lambda b, k, l: s.kf_rotation(l, k), "edge_blank")
    ctx.close()
    assert ok


def test_sequence_feature_age_culling():
    """max_feature_age lowered to 3 keyframes so that the age limit (Tracker.cpp:150-160) removes tracks
    inside a 26-frame sequence (ages only advance at keyframes; the shipped limit of 25 needs > 100 frames)."""
    import dataclasses
    from kimera_vio_b200.params import FrontendParams
    params = dataclasses.replace(FrontendParams.euroc(), max_feature_track_age=3)
    p, rig, ctx = H.euroc_setup(batch=1, params=params)
    orig = StereoRig(CameraParams.euroc_left(), CameraParams.euroc_right())
    s, fr = H.synth_frames(26, seed=31337)
    frames = [(f.left, f.right, f.timestamp) for f in fr]
    fe = ofe.StereoFrontend(p, orig)
    ok = run_sequence(ctx, [fe], [frames], lambda b, k, l: s.kf_rotation(l, k), "edge_age")
    ctx.close()
    assert ok
