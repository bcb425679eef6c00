"""Parity on the larger BASELINE.json configurations (parity-test cases, not bench lines):
C3-like 1280x720 / 500 features, C4-like 1920x1080 / 1000 features and the C5 image size
3840x2160 / 2000 features on ONE GPU (the 8-GPU row tiling of C5 is not built), synthetic stereo, a
few frames of the whole front-end against the oracle."""
import dataclasses

import numpy as np
import pytest

import helpers as H
from kimera_vio_b200 import lib as kl
from kimera_vio_b200.params import CameraParams, FrontendParams
from kimera_vio_b200.rig import StereoRigSetup
from kimera_vio_b200.synth import SynthStream
from oracle import frontend as ofe
from oracle.rig import StereoRig
from test_gpu_sequence import run_sequence

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("w,h,feats,frames", [(1280, 720, 500, 6), (1920, 1080, 1000, 5), (3840, 2160, 2000, 3)])
def test_sequence_larger_configs(w, h, feats, frames):
    left, right = CameraParams.euroc_left().scaled(w, h), CameraParams.euroc_right().scaled(w, h)
    p = dataclasses.replace(FrontendParams.euroc(), max_features_per_frame=feats)
    rig = StereoRigSetup(left, right)
    cfg = kl.make_config(p, w, h, batch=1, sobel_cpu_tail_start=H.sobel_cpu_tail_start(w))
    ctx = kl.Context(cfg, rig.to_c())
    s = SynthStream(left, right, rig.R1, seed=77)
    fr = [s.frame(k) for k in range(frames)]
    orig = StereoRig(left, right)
    fe = ofe.StereoFrontend(p, orig)
    ok = run_sequence(ctx, [fe], [[(f.left, f.right, f.timestamp) for f in fr]],
                      lambda b, k, l: s.kf_rotation(l, k), "cfg%dx%d" % (w, h))
    ctx.close()
    assert ok


@pytest.mark.parametrize("rig_name,frames", [("uHumans2", 12), ("D455", 10)])
def test_sequence_shipped_rigs(rig_name, frames):
    """The reference's own params/uHumans2 (720x480, no distortion, maxFeatureAge 15, min_distance 8,
    max_disparity_since_lkf 200) and params/D455 (640x480 radtan, 5-point + 3-point RANSAC) parameter and
    camera files -- values from tests/golden/rigs.json -- on a synthetic stream rendered with that rig."""
    p, left, right = H.shipped_rig(rig_name)
    rig = StereoRigSetup(left, right)
    cfg = kl.make_config(p, rig.W, rig.H, batch=1, sobel_cpu_tail_start=H.sobel_cpu_tail_start(rig.W))
    ctx = kl.Context(cfg, rig.to_c())
    s = SynthStream(left, right, rig.R1, seed=4242)
    fr = [s.frame(k) for k in range(frames)]
    fe = ofe.StereoFrontend(p, StereoRig(left, right))
    ok = run_sequence(ctx, [fe], [[(f.left, f.right, f.timestamp) for f in fr]],
                      lambda b, k, l: s.kf_rotation(l, k), "rig_" + rig_name)
    ctx.close()
    assert ok
