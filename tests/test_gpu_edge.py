"""Edge-case sequences through the device-resident FSM vs the oracle: a textureless stereo pair in the
middle of a sequence (tracking into it, detection on it, every track lost on the next frame ->
re-detection without a keyframe, StereoVisionImuFrontend.cpp:312-323), a sensor-noise-only pair, and
a sequence long enough for max_feature_age culling (Tracker.cpp:150-160).
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


def test_sequence_blank_and_noise_frames():
    p, rig, ctx = H.euroc_setup(batch=1)
    orig = StereoRig(CameraParams.euroc_left(), CameraParams.euroc_right())
    s, fr = H.synth_frames(12, seed=9090)
    frames = [(f.left.copy(), f.right.copy(), f.timestamp) for f in fr]
    blank = np.full_like(frames[0][0], 128)
    rng = np.random.default_rng(5)
    noise = np.clip(110 + rng.normal(0, 2.0, blank.shape), 0, 255).astype(np.uint8)
    frames[4] = (blank, blank.copy(), frames[4][2])          # textureless pair
    frames[8] = (noise, noise.copy(), frames[8][2])          # sensor noise only
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
