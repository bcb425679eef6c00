"""Row f2 (partial): MonoVisionImuFrontend (src/frontend/MonoVisionImuFrontend.cpp:194-371) -- the stereo kernels without
the stereo half, frontend_type = 1 -- against oracle/mono.py on a synthetic stream: keypoints, landmark ids, ages,
bearing vectors (no rectification rotation), keyframe decisions, statuses (mono reset every frame, stereo DISABLED),
keypoints_undistorted_ (P = K, R = I) and the smart mono measurements (uR = NaN)."""
import numpy as np
import pytest

import helpers as H
from kimera_vio_b200 import lib as kl
from kimera_vio_b200.params import CameraParams, FrontendParams
from kimera_vio_b200.rig import MonoRigSetup
from oracle import frontend as ofe
from oracle.mono import MonoFrontend

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("identity_rotation", [False, True])
def test_mono_frontend_sequence(identity_rotation):
    N = 14
    p = FrontendParams.euroc()
    cam = CameraParams.euroc_left()
    rig = MonoRigSetup(cam)
    cfg = kl.make_config(p, rig.W, rig.H, batch=1, sobel_cpu_tail_start=H.sobel_cpu_tail_start(rig.W), mono=True)
    ctx = kl.Context(cfg, rig.to_c())
    s, fr = H.synth_frames(N, seed=20240)
    fe = MonoFrontend(p, cam)
    lkf, bad, n_kf = 0, [], 0
    for k, f in enumerate(fr):
        R = np.eye(3) if identity_rotation else s.kf_rotation(lkf, k)
        pk = ctx.step([f.left], [f.left], [f.timestamp], np.array([R]))[0]
        o, is_kf, smart = fe.spin(ofe.Frame(k, f.timestamp, f.left, cam), R)
        rec = dict(k=k, n=(int(pk["n"]), len(o.keypoints)), kf=(int(pk["is_keyframe"]), int(is_kf)))
        ok = pk["n"] == len(o.keypoints) and bool(pk["is_keyframe"]) == bool(is_kf)
        if ok and pk["n"]:
            kp = np.array(o.keypoints, np.float32).reshape(-1, 2)
            g = np.stack([pk["kp_x"], pk["kp_y"]], 1)
            rec["kp_err"] = float(np.abs(g - kp).max())
            ok &= rec["kp_err"] <= 1e-3
            ok &= np.array_equal(pk["landmark"], np.array(o.landmarks, np.int64)) and np.array_equal(pk["age"], np.array(o.landmarks_age))
            rec["versor_err"] = float(np.abs(pk["versor"] - np.array(o.versors).reshape(-1, 3)).max())
            ok &= rec["versor_err"] < 1e-5
        rec["status"] = (int(pk["mono_status"]), int(fe.mono_status), int(pk["stereo_status"]))
        ok &= pk["mono_status"] == fe.mono_status and pk["stereo_status"] == ofe.DISABLED or k == 0
        if ok and is_kf:
            n_kf += 1
            us = np.array([st for st, _ in o.keypoints_undistorted], np.int32)
            ux = np.array([q for _, q in o.keypoints_undistorted], np.float32).reshape(-1, 2)
            ok &= np.array_equal(pk["left_status"], us)
            rec["undist_err"] = float(np.abs(np.stack([pk["left_rect_x"], pk["left_rect_y"]], 1) - ux).max())
            ok &= rec["undist_err"] <= 2e-3
            ok &= (pk["right_status"] == -1).all()
            if k > 0:
                ok &= pk["n_smart"] == len(smart)
                if pk["n_smart"] == len(smart) and smart:
                    ok &= np.array_equal(pk["smart_lmk"], np.array([m[0] for m in smart], np.int64))
                    ok &= np.isnan(pk["smart_uR"]).all()
                    ok &= np.abs(pk["smart_uL"] - np.array([m[1] for m in smart])).max() <= 2e-3
        rec["ok"] = bool(ok)
        H.diag("mono_sequence", **rec)
        if not ok:
            bad.append(rec)
        if is_kf:
            lkf = k
    ctx.close()
    assert not bad, bad[:3]
    assert n_kf >= 3
