"""Row f3 (the step before the path): cv::equalizeHist as UtilsOpenCV::ReadAndConvertToGrayScale applies it when
stereo_matching_params.equalize_image is set -- bit-exact against cv2.equalizeHist, and a whole sequence with the
flag on against the oracle fed pre-equalized images; keys that would change results are rejected, not ignored."""
import dataclasses

import cv2
import numpy as np
import pytest

import helpers as H
from kimera_vio_b200 import lib as kl
from kimera_vio_b200.params import CameraParams, FrontendParams
from oracle import frontend as ofe
from oracle.rig import StereoRig
from test_gpu_sequence import run_sequence

pytestmark = pytest.mark.gpu


def test_equalize_hist_bit_exact():
    p, rig, ctx = H.euroc_setup(batch=1)
    g, lefts, rights = H.golden()
    rng = np.random.default_rng(0)
    imgs = [lefts[0], rights[1], np.full((480, 752), 77, np.uint8),
            (rng.integers(0, 2, (480, 752)) * 200 + 20).astype(np.uint8),
            rng.integers(0, 256, (480, 752)).astype(np.uint8),
            np.clip(rng.normal(128, 3, (480, 752)), 0, 255).astype(np.uint8)]
    for k, im in enumerate(imgs):
        got = ctx.equalize_hist(im)
        want = cv2.equalizeHist(im)
        assert np.array_equal(got, want), k
    ctx.close()


def test_sequence_with_equalize_image():
    N = 6
    p = dataclasses.replace(FrontendParams.euroc(), equalize_image=True)
    p2, rig, ctx = H.euroc_setup(batch=1, params=p)
    orig = StereoRig(CameraParams.euroc_left(), CameraParams.euroc_right())
    s, fr = H.synth_frames(N, seed=20240)
    fe = ofe.StereoFrontend(FrontendParams.euroc(), orig)

    class Eq:     # the oracle sees what the reference's data provider would hand over
        def __init__(self, f):
            self.f = f
    lkf, ok = 0, True
    from test_gpu_sequence import compare_packet, packet_ok
    for k, f in enumerate(fr):
        R = s.kf_rotation(lkf, k)
        pk = ctx.step([f.left], [f.right], [f.timestamp], np.array([R]))[0]
        o = fe.spin(ofe.StereoFrame.make(k, f.timestamp, cv2.equalizeHist(f.left), cv2.equalizeHist(f.right), orig), R)
        rec = compare_packet("equalize/f%d" % k, pk, o)
        rec["ok"] = packet_ok(rec)
        H.diag("sequence", **rec)
        ok &= rec["ok"]
        if o.is_keyframe:
            lkf = k
    ctx.close()
    assert ok


def test_unsupported_keys_are_rejected():
    rig = H.StereoRigSetup(CameraParams.euroc_left(), CameraParams.euroc_right())
    for key in ("optimize_2d2d_pose_from_inliers", "optimize_3d3d_pose_from_inliers"):
        p = dataclasses.replace(FrontendParams.euroc(), **{key: True})
        cfg = kl.make_config(p, rig.W, rig.H, batch=1)
        with pytest.raises(kl.KvfeError):
            kl.Context(cfg, rig.to_c())
