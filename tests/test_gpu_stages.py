"""GPU parity tests, stage by stage, through the C-ABI (ctypes -> libkvfe.so) against the oracle
(cv2 4.13 / oracle/*.py).  Bit-exact for images, maps, pyramids, response maps, corner indices,
status flags and inlier masks; <= 1e-3 px for LK / sub-pixel coordinates (north_star tolerance)."""
import zlib

import cv2
import numpy as np
import pytest

import helpers as H
import scenes
from kimera_vio_b200 import lib as kl
from kimera_vio_b200.params import CameraParams, FrontendParams
from kimera_vio_b200.rig import StereoRigSetup
from oracle import frontend as ofe
from oracle import ransac as ors
from oracle.rig import StereoRig

pytestmark = pytest.mark.gpu

TOL_PX = 1e-3


@pytest.fixture(scope="module")
def env():
    p, rig, ctx = H.euroc_setup(batch=1)
    orig = StereoRig(CameraParams.euroc_left(), CameraParams.euroc_right())
    g, lefts, rights = H.golden()
    s, frames = H.synth_frames(3)
    yield dict(p=p, rig=rig, ctx=ctx, orig=orig, g=g, lefts=lefts, rights=rights, synth=frames, stream=s)
    ctx.close()


def _images(env):
    return [("euroc", env["lefts"][0], env["rights"][0]), ("synth", env["synth"][0].left, env["synth"][0].right)]


def test_maps_bit_exact(env):
    o = env["orig"]
    for cam, (mx, my) in enumerate(((o.map_lx, o.map_ly), (o.map_rx, o.map_ry))):
        gx, gy = env["ctx"].rectify_maps(cam)
        bad = int((gx != mx).sum() + (gy != my).sum())
        H.diag("maps", cam=cam, mismatches=bad)
        assert bad == 0


def test_rectify_bit_exact(env):
    o = env["orig"]
    for name, L, R in _images(env):
        gl, gr = env["ctx"].rectify_pair(L, R)
        el, er = o.rectify_left(L), o.rectify_right(R)
        bad = int((gl != el).sum() + (gr != er).sum())
        H.diag("rectify", image=name, mismatches=bad, total=int(el.size * 2))
        assert bad == 0
    # committed golden (oracle run in the build container)
    gl, gr = env["ctx"].rectify_pair(env["lefts"][0], env["rights"][0])
    assert zlib.crc32(gl.tobytes()) == int(env["g"]["rect_left_crc_0"][0])
    assert zlib.crc32(gr.tobytes()) == int(env["g"]["rect_right_crc_0"][0])


def test_pyramid_bit_exact(env):
    for name, L, _ in _images(env):
        lv = env["ctx"].pyramid(L)
        _, ref = cv2.buildOpticalFlowPyramid(L, (24, 24), 4, withDerivatives=False)
        assert len(lv) == 4
        prev = L
        for i, g in enumerate(lv):
            e = cv2.pyrDown(prev)
            bad = int((g != e).sum())
            H.diag("pyramid", image=name, level=i + 1, mismatches=bad)
            assert bad == 0
            prev = e


def test_min_eigen_response_bit_exact(env):
    for name, L, _ in _images(env):
        g = env["ctx"].min_eigen_response(L)
        e = cv2.cornerMinEigenVal(L, 3, ksize=3)
        bad = g != e
        H.diag("mineig", image=name, mismatches=int(bad.sum()),
               max_abs=float(np.abs(g - e).max()), first=np.argwhere(bad)[:5])
        assert int(bad.sum()) == 0
    g = env["ctx"].min_eigen_response(env["lefts"][0])
    H.diag("mineig_golden", crc_gpu=zlib.crc32(g.tobytes()), crc_golden=int(env["g"]["eig_crc_0"][0]))


def _oracle_raw(det, img, kps, lmks, cam):
    fr = ofe.Frame(0, 0, img, cam, keypoints=list(kps), landmarks=list(lmks))
    return det.raw_feature_detection(img, det.build_mask(fr))


def test_gftt_raw_bit_exact(env):
    det = ofe.FeatureDetector(env["p"])
    cam = env["orig"].left
    rng = np.random.default_rng(1)
    for name, L, _ in _images(env):
        for masked in (False, True):
            kps, lmks = [], []
            if masked:
                kps = [(np.float32(x), np.float32(y)) for x, y in
                       zip(rng.uniform(0, 752, 150), rng.uniform(0, 480, 150))]
                kps += [(np.float32(20.5), np.float32(21.5)), (np.float32(3.2), np.float32(470.7))]
                lmks = [int(i) if i % 7 else -1 for i in range(len(kps))]
            raw = _oracle_raw(det, L, kps, lmks, cam)
            e = np.array([k.pt for k in raw], np.float32).reshape(-1, 2)
            g, resp = env["ctx"].detect_raw(L, kps, lmks)
            same = g.shape == e.shape and bool(np.all(g == e))
            H.diag("gftt_raw", image=name, masked=masked, n_gpu=len(g), n_ref=len(e), same=same)
            assert g.shape == e.shape
            assert np.array_equal(g, e)
    g, _ = env["ctx"].detect_raw(env["lefts"][0])
    assert np.array_equal(g, env["g"]["gftt_raw_0"])


def test_detect_with_anms_and_subpix(env):
    det = ofe.FeatureDetector(env["p"])
    cam = env["orig"].left
    rng = np.random.default_rng(2)
    for name, L, _ in _images(env):
        for masked, need in ((False, 300), (True, 120), (True, 3000)):
            kps, lmks = [], []
            if masked:
                kps = [(np.float32(x), np.float32(y)) for x, y in
                       zip(rng.uniform(0, 752, 180), rng.uniform(0, 480, 180))]
                lmks = list(range(len(kps)))
            fr = ofe.Frame(0, 0, L, cam, keypoints=list(kps), landmarks=list(lmks))
            e = det.detect_corners(fr, need)
            g = env["ctx"].detect(L, kps, lmks, need)
            n_ok = len(g) == len(e)
            err = float(np.abs(g - e).max()) if n_ok and len(e) else -1.0
            n_bad = int((np.abs(g - e).max(axis=1) > TOL_PX).sum()) if n_ok and len(e) else -1
            H.diag("detect", image=name, masked=masked, need=need, n_gpu=len(g), n_ref=len(e),
                   max_err=err, n_over_tol=n_bad)
            assert n_ok
            # integer corner identity (pre-subpix) is implied by <=1e-3 agreement of the refined
            # positions; refined positions must agree within the north_star tolerance
            assert n_bad == 0, "sub-pixel corners differ by more than 1e-3 px (max %.3g)" % err
    g = env["ctx"].detect(env["lefts"][0], [], [], 300)
    assert len(g) == len(env["g"]["detect_0"])
    assert np.abs(g - env["g"]["detect_0"]).max() <= TOL_PX


def test_lk_tracking(env):
    p = env["p"]
    pairs = [("euroc", env["lefts"][0], env["lefts"][1]), ("euroc_gap", env["lefts"][0], env["lefts"][4]),
             ("synth", env["synth"][0].left, env["synth"][1].left)]
    rng = np.random.default_rng(3)
    for name, A, B in pairs:
        c = cv2.goodFeaturesToTrack(A, 300, 0.001, 20).reshape(-1, 2)
        # add border / textureless / off-grid points
        extra = np.array([[2.5, 3.5], [749.2, 477.1], [375.3, 1.2], [1.1, 240.9], [700.7, 10.2]], np.float32)
        pts = np.concatenate([c, extra, c[:40] + rng.uniform(-0.5, 0.5, (40, 2)).astype(np.float32)]).astype(np.float32)
        for rot in (np.eye(3), scenes.expmap([0.004, -0.003, 0.002])):
            pred = ofe.predict_sparse_flow([tuple(q) for q in pts], rot, env["orig"].left.K, 752, 480,
                                           p.optical_flow_predictor_type)
            a = pts.reshape(-1, 1, 2)
            b = np.array(pred, np.float32).reshape(-1, 1, 2)
            crit = (cv2.TERM_CRITERIA_COUNT + cv2.TERM_CRITERIA_EPS, p.klt_max_iter, p.klt_eps)
            nxt, st, _ = cv2.calcOpticalFlowPyrLK(A, B, a, b.copy(), winSize=(24, 24), maxLevel=p.klt_max_level,
                                                  criteria=crit, flags=cv2.OPTFLOW_USE_INITIAL_FLOW)
            gp, gn, gs = env["ctx"].track(A, B, rot, pts)
            pred_bad = int((gp != np.array(pred, np.float32)).sum())
            st_bad = int((gs != st.reshape(-1)).sum())
            ok = st.reshape(-1) == 1
            d = np.abs(gn - nxt.reshape(-1, 2))[ok & (gs == 1)]
            H.diag("lk", pair=name, rot=not np.allclose(rot, np.eye(3)), n=len(pts), pred_mismatch=pred_bad,
                   status_mismatch=st_bad, n_fail_ref=int((~ok).sum()), max_err=float(d.max()) if len(d) else 0.0,
                   n_over_tol=int((d.max(axis=1) > TOL_PX).sum()) if len(d) else 0,
                   n_exact=int((d.max(axis=1) == 0).sum()) if len(d) else 0)
            assert pred_bad == 0
            assert st_bad == 0
            assert len(d) == 0 or d.max() <= TOL_PX


def test_undistort_and_bearing(env):
    o = env["orig"]
    rng = np.random.default_rng(4)
    pts = np.stack([rng.uniform(0, 752, 500), rng.uniform(0, 480, 500)], 1).astype(np.float32)
    for cam, camp, R, P in ((0, o.left, o.R1, o.P1), (1, o.right, o.R2, o.P2)):
        for useR, useP in ((False, False), (True, False), (True, True)):
            e = cv2.undistortPoints(pts.reshape(-1, 1, 2), camp.K, camp.D, R=R if useR else None,
                                    P=P if useP else None).reshape(-1, 2)
            g = env["ctx"].undistort_keypoints(cam, useR, useP, pts)
            bad = int((g != e).sum())
            H.diag("undistort", cam=cam, useR=useR, useP=useP, mismatches=bad, max_abs=float(np.abs(g - e).max()))
            assert bad == 0
    v = env["ctx"].bearing_vectors(pts)
    e = np.array(ofe.get_bearing_vectors([tuple(q) for q in pts], o.left, o.R1))
    H.diag("bearing", mismatches=int((v != e).sum()), max_abs=float(np.abs(v - e).max()))
    assert np.array_equal(v, e)


def test_sparse_stereo(env):
    p, o = env["p"], env["orig"]
    m = ofe.StereoMatcher(p, o)
    for name, L, R in _images(env):
        c = cv2.goodFeaturesToTrack(L, 300, 0.001, 20).reshape(-1, 2).astype(np.float32)
        extra = np.array([[5.2, 4.1], [745.0, 3.0], [3.0, 476.0], [748.9, 478.2], [30.5, 240.5], [720.5, 200.5]], np.float32)
        kps = np.concatenate([c, extra])
        sf = ofe.StereoFrame.make(0, 0, L, R, o)
        sf.left_frame.keypoints = [(np.float32(x), np.float32(y)) for x, y in kps]
        sf.left_frame.versors = ofe.get_bearing_vectors(sf.left_frame.keypoints, o.left, o.R1)
        m.sparse_stereo_reconstruction(sf)
        g = env["ctx"].sparse_stereo(L, R, kps, np.array(sf.left_frame.versors))
        els = np.array([s for s, _ in sf.left_keypoints_rectified])
        elx = np.array([q for _, q in sf.left_keypoints_rectified], np.float32)
        ers = np.array([s for s, _ in sf.right_keypoints_rectified])
        erx = np.array([q for _, q in sf.right_keypoints_rectified], np.float32)
        rec = dict(image=name, n=len(kps),
                   left_status_mismatch=int((g["left_status"] != els).sum()),
                   left_xy_mismatch=int((np.stack([g["left_rect_x"], g["left_rect_y"]], 1) != elx).sum()),
                   right_status_mismatch=int((g["right_status"] != ers).sum()),
                   right_xy_mismatch=int((np.stack([g["right_rect_x"], g["right_rect_y"]], 1) != erx).sum()),
                   depth_max_err=float(np.abs(g["depth"] - np.array(sf.keypoints_depth)).max()),
                   p3d_max_err=float(np.abs(g["points_3d"] - np.array(sf.keypoints_3d)).max()),
                   right_raw_mismatch=int((np.stack([g["right_x"], g["right_y"]], 1) !=
                                           np.array(sf.right_frame.keypoints, np.float32)).sum()),
                   rect_mismatch=int((g["left_rect"] != sf.left_img_rectified).sum() +
                                     (g["right_rect"] != sf.right_img_rectified).sum()),
                   n_valid=int((ers == 0).sum()))
        H.diag("sparse_stereo", **rec)
        assert rec["left_status_mismatch"] == 0 and rec["left_xy_mismatch"] == 0
        assert rec["right_status_mismatch"] == 0
        assert rec["right_xy_mismatch"] == 0, "disparity differs (float-DFT near-tie in cv2.matchTemplate?)"
        assert rec["depth_max_err"] == 0 and rec["p3d_max_err"] == 0 and rec["right_raw_mismatch"] == 0
        assert rec["rect_mismatch"] == 0


def test_sparse_stereo_near_tie_divergence_is_counted(env):
    """cv::matchTemplate(TM_SQDIFF) evaluates the squared difference through a float DFT (absolute error up
    to ~32 at magnitudes of 5e6..8e6), the GPU matcher in exact int32.  The arg-min can only
    differ where two shifts are tied to within that error.  On textured input the divergence count is asserted
    ZERO (here and over whole sequences in test_gpu_long.py); on a sensor-noise-only pair -- hundreds of nearly
    tied shifts per keypoint, the worst case -- it is COUNTED, reported and bounded."""
    p, o, ctx = env["p"], env["orig"], env["ctx"]
    m = ofe.StereoMatcher(p, o)
    rng = np.random.default_rng(5)
    noise = np.clip(110 + rng.normal(0, 2.0, env["lefts"][0].shape), 0, 255).astype(np.uint8)
    cases = [("euroc", env["lefts"][1], env["rights"][1], 0), ("synth", env["synth"][1].left, env["synth"][1].right, 0),
             ("noise_only", noise, noise.copy(), None)]
    for name, L, R, allowed in cases:
        c = cv2.goodFeaturesToTrack(L, 300, 0.001, 20).reshape(-1, 2).astype(np.float32)
        sf = ofe.StereoFrame.make(0, 0, L, R, o)
        sf.left_frame.keypoints = [(np.float32(x), np.float32(y)) for x, y in c]
        sf.left_frame.versors = ofe.get_bearing_vectors(sf.left_frame.keypoints, o.left, o.R1)
        m.sparse_stereo_reconstruction(sf)
        g = ctx.sparse_stereo(L, R, c, np.array(sf.left_frame.versors))
        ers = np.array([s for s, _ in sf.right_keypoints_rectified])
        erx = np.array([q for _, q in sf.right_keypoints_rectified], np.float32).reshape(-1, 2)
        gx = np.stack([g["right_rect_x"], g["right_rect_y"]], 1)
        matched = (ers == 0) | (g["right_status"] == 0)
        div = int((((gx != erx).any(axis=1)) & matched).sum() + (g["right_status"] != ers).sum())
        H.diag("stereo_near_tie", image=name, keypoints=len(c), matched=int(matched.sum()), divergent=div)
        if allowed is not None:
            assert div == allowed, (name, div)
        else:
            assert div <= max(3, len(c) // 25), (name, div, len(c))     # measured in round 1: 2 of 286


@pytest.mark.parametrize("variant", ["subpixel", "extra_rows", "templ_51x7", "near_range"])
def test_sparse_stereo_param_variants(env, variant):
    """Stereo-matcher parameter variants (StereoMatchingParams.h:46-60): cornerSubPix on the right match,
    a stripe taller than the template (several row shifts), a smaller template (other GEMM tiling), a
    shorter stripe (min_point_dist)."""
    import dataclasses
    kw = {"subpixel": dict(subpixel_refinement_stereo=True), "extra_rows": dict(stripe_extra_rows=2),
          "templ_51x7": dict(templ_cols=51, templ_rows=7), "near_range": dict(min_point_dist=1.0)}[variant]
    params = dataclasses.replace(env["p"], **kw)
    p, rig, ctx = H.euroc_setup(batch=1, params=params)
    o = env["orig"]
    m = ofe.StereoMatcher(p, o)
    name, L, R = _images(env)[1]
    c = cv2.goodFeaturesToTrack(L, 200, 0.001, 20).reshape(-1, 2).astype(np.float32)
    extra = np.array([[5.2, 4.1], [745.0, 3.0], [3.0, 476.0], [748.9, 478.2], [30.5, 240.5], [720.5, 200.5]], np.float32)
    kps = np.concatenate([c, extra])
    sf = ofe.StereoFrame.make(0, 0, L, R, o)
    sf.left_frame.keypoints = [(np.float32(x), np.float32(y)) for x, y in kps]
    sf.left_frame.versors = ofe.get_bearing_vectors(sf.left_frame.keypoints, o.left, o.R1)
    m.sparse_stereo_reconstruction(sf)
    g = ctx.sparse_stereo(L, R, kps, np.array(sf.left_frame.versors))
    ers = np.array([s_ for s_, _ in sf.right_keypoints_rectified])
    erx = np.array([q for _, q in sf.right_keypoints_rectified], np.float32)
    gx = np.stack([g["right_rect_x"], g["right_rect_y"]], 1)
    rec = dict(variant=variant, n=len(kps), right_status_mismatch=int((g["right_status"] != ers).sum()),
               right_xy_max_err=float(np.abs(gx - erx).max()), right_xy_mismatch=int((gx != erx).any(axis=1).sum()),
               depth_max_rel=float(np.max(np.abs(g["depth"] - np.array(sf.keypoints_depth)) / np.maximum(1e-12, np.abs(np.array(sf.keypoints_depth))))),
               n_valid=int((ers == 0).sum()))
    H.diag("sparse_stereo_variant", **rec)
    ctx.close()
    assert rec["right_status_mismatch"] == 0
    if variant == "subpixel":
        assert rec["right_xy_max_err"] <= 1e-3          # sub-pixel coordinates: tolerance of north_star
    else:
        assert rec["right_xy_mismatch"] == 0


def test_ransac_2pt(env):
    cam = CameraParams.euroc_left()
    for planar, n_in, n_out in ((False, 80, 0), (False, 80, 20), (True, 80, 20), (False, 200, 100), (False, 1, 0)):
        for R, T in ((np.eye(3), np.array([1.0, 0, 0])), (scenes.expmap([0.01, -0.02, 0.015]), np.array([0.3, 0.1, -0.05]))):
            rng = np.random.default_rng(7)
            f_ref, f_cur = scenes.mono_scene(rng, cam, R, T, n_in, n_out, planar)
            prob = ors.Problem2d2dGivenRot(f_ref, f_cur, R, ors.rnd_table(4096))
            ok, pose, inl = ors.run_ransac(prob, 1e-6, 100, 0.995)
            st, gpose, ginl = env["ctx"].ransac_mono(f_ref, f_cur, R)
            est = ors.INVALID if not ok else (ors.FEW_MATCHES if len(inl) < 10 else ors.VALID)
            H.diag("ransac_2pt", planar=planar, n_in=n_in, n_out=n_out, status_gpu=st, status_ref=est,
                   inliers_equal=ginl == inl, n_inl_gpu=len(ginl), n_inl_ref=len(inl),
                   pose_err=float(np.abs(gpose - pose).max()))
            assert st == est
            assert ginl == inl
            assert np.abs(gpose - pose).max() < 1e-9


def test_ransac_3pt(env):
    rig = env["orig"]
    R, T = scenes.expmap([0.1, 0.1, 0.1]), np.array([rig.baseline, 0, 0])
    for n_in, n_out in ((3, 0), (40, 0), (80, 40), (2, 0)):
        rng = np.random.default_rng(11)
        sc = scenes.stereo_scene(rng, rig, R, T, n_in, n_out, [rig.baseline * 10, rig.baseline * 20])
        prob = ors.Problem3d3d(sc["p_ref"], sc["p_cur"], ors.rnd_table(4096))
        ok, pose, inl = ors.run_ransac(prob, 1.0, 100, 0.995)
        st, gpose, ginl = env["ctx"].ransac_stereo_3pt(sc["p_ref"], sc["p_cur"])
        est = ors.INVALID if not ok else (ors.FEW_MATCHES if len(inl) < 5 else ors.VALID)
        H.diag("ransac_3pt", n_in=n_in, n_out=n_out, status_gpu=st, status_ref=est, inliers_equal=ginl == inl,
               pose_err=float(np.abs(gpose - pose).max()))
        assert st == est
        assert ginl == inl
        assert np.abs(gpose - pose).max() < 1e-6


def test_ransac_1pt(env):
    rig = env["orig"]
    R, T = scenes.expmap([0.1, 0.1, 0.1]), np.array([rig.baseline, 0, 0])
    calib = (rig.fx, rig.fy, rig.cx, rig.cy, rig.baseline)
    for n_in, n_out in ((3, 0), (40, 0), (80, 40), (250, 60), (1, 0)):
        rng = np.random.default_rng(13)
        sc = scenes.stereo_scene(rng, rig, R, T, n_in, n_out, [rig.baseline * 10, rig.baseline * 20])
        R1 = rig.R1
        p_ref, p_cur = (R1 @ sc["p_ref"].T).T, (R1 @ sc["p_cur"].T).T
        Rr = R1 @ R @ R1.T
        matches = [(i, i) for i in range(n_in + n_out)]
        est, pose, inl, info = ors.outlier_rejection_3d3d_given_rotation(
            sc["ref_left"], sc["ref_right"], sc["cur_left"], sc["cur_right"], p_ref, p_cur, calib, matches, Rr, 1.0, 5)
        st, gpose, ginl, ginfo = env["ctx"].ransac_stereo_1pt(sc["ref_left"], sc["ref_right"], sc["cur_left"],
                                                            sc["cur_right"], p_ref, p_cur, Rr)
        H.diag("ransac_1pt", n_in=n_in, n_out=n_out, status_gpu=st, status_ref=est, inliers_equal=ginl == inl,
               n_inl=len(inl), pose_err=float(np.abs(gpose - pose).max()),
               info_rel_err=float(np.abs(ginfo - info).max() / (np.abs(info).max() + 1e-30)))
        assert st == est
        assert ginl == inl
        assert np.abs(gpose - pose).max() < 1e-6
        assert np.abs(ginfo - info).max() <= 1e-9 * (np.abs(info).max() + 1e-30)


def test_ransac_5pt_nister():
    """5-point mono RANSAC (ransac_use_2point_mono = 0, e.g. params/D455): inlier masks and statuses
    against the oracle on the reference's synthetic scenes (tests/testTracker.cpp:704-801)."""
    import dataclasses
    p = dataclasses.replace(FrontendParams.euroc(), ransac_use_2point_mono=False, ransac_use_1point_stereo=False,
                            ransac_max_iterations=1000)
    p2, rig, ctx = H.euroc_setup(batch=1, params=p)
    cam = CameraParams.euroc_left()
    R, T = scenes.expmap([0.01, 0.01, 0.01]), np.array([1.0, 0, 0])
    for planar, n_in, n_out in ((False, 82, 0), (False, 80, 40), (True, 80, 40), (False, 7, 0)):
        rng = np.random.default_rng(3)
        f_ref, f_cur = scenes.mono_scene(rng, cam, R, T, n_in, n_out, planar)
        prob = ors.Problem2d2dNister(f_ref, f_cur, ors.rnd_table(32768)) if n_in + n_out >= 8 else None
        if prob is not None:
            ok, pose, inl = ors.run_ransac(prob, 1e-6, 1000, 0.995)
        else:
            ok, pose, inl = False, None, []
        st, gpose, ginl = ctx.ransac_mono(f_ref, f_cur, None)
        est = ors.INVALID if not ok else (ors.FEW_MATCHES if len(inl) < 10 else ors.VALID)
        H.diag("ransac_5pt", planar=planar, n_in=n_in, n_out=n_out, status_gpu=st, status_ref=est,
               inliers_equal=ginl == inl, n_inl_gpu=len(ginl), n_inl_ref=len(inl),
               pose_err=float(np.abs(gpose - pose).max()) if ok and st != ors.INVALID else -1.0)
        assert st == est
        assert ginl == inl
        if ok:
            assert np.allclose(gpose[:, :3], R, atol=1e-3)
            t = gpose[:, 3] / np.linalg.norm(gpose[:, 3])
            assert np.allclose(t, T, atol=1e-3)
    ctx.close()


@pytest.mark.parametrize("order", ["gcc", "clang"])
def test_ransac_reference_seeded_scenes(order):
    """The four RANSAC kernels on the scenes of the reference's own tests regenerated from the reference's RNG streams
    (tests/golden/tracker_scenes.npz): inlier lists equal to the oracle's, which satisfy the reference's assertions
    (tests/test_oracle_ransac.py::test_reference_seeded_scenes)."""
    import dataclasses
    from test_oracle_ransac import seeded_scenes
    z, left, right = seeded_scenes()
    rig = StereoRigSetup(left, right)

    def make(**kw):
        p = dataclasses.replace(FrontendParams.euroc(), **kw)
        cfg = kl.make_config(p, rig.W, rig.H, batch=1, sobel_cpu_tail_start=H.sobel_cpu_tail_start(rig.W))
        return kl.Context(cfg, rig.to_c())
    bad = []
    ctx = make(ransac_use_2point_mono=False, ransac_use_1point_stereo=False, ransac_max_iterations=1000)
    for ci in range(3):
        pre = "%s/5pt/%d/" % (order, ci)
        st, gpose, ginl = ctx.ransac_mono(z[pre + "f_ref"], z[pre + "f_cur"], None)
        want = [int(v) for v in z[pre + "oracle_inliers"]]
        # the reference asserts the status and the landmark sets only (testTracker.cpp:782-797); on the planar scene
        # several hypotheses share the inlier set and hypothesis-level parity with OpenGV is not claimed
        if not (st == ors.VALID and ginl == want and (ci == 2 or np.allclose(gpose[:, :3], z["R_5pt"], atol=1e-3))):
            bad.append((pre, st, len(ginl), len(want)))
    ctx.close()
    ctx = make(ransac_threshold_stereo=0.3)
    for ci in range(3):
        pre = "%s/2pt/%d/" % (order, ci)
        st, gpose, ginl = ctx.ransac_mono(z[pre + "f_ref"], z[pre + "f_cur"], np.eye(3))
        want = [int(v) for v in z[pre + "oracle_inliers"]]
        if not (st == ors.VALID and ginl == want and np.abs(gpose - z[pre + "oracle_pose"]).max() < 1e-9):
            bad.append((pre, st, len(ginl), len(want)))
    for ci in range(4):
        pre = "%s/3pt/%d/" % (order, ci)
        st, gpose, ginl = ctx.ransac_stereo_3pt(z[pre + "p_ref"], z[pre + "p_cur"])
        want = [int(v) for v in z[pre + "oracle_inliers"]]
        if not (ginl == want and st == (ors.VALID if len(want) >= 5 else ors.FEW_MATCHES) and np.abs(gpose - z[pre + "oracle_pose"]).max() < 1e-6):
            bad.append((pre, st, len(ginl), len(want)))
    ctx.close()
    ctx = make()
    R1 = np.asarray(z["R1"])
    for ci in range(4):
        pre = "%s/1pt/%d/" % (order, ci)
        p_ref, p_cur = (R1 @ z[pre + "p_ref"].T).T, (R1 @ z[pre + "p_cur"].T).T
        st, gpose, ginl, ginfo = ctx.ransac_stereo_1pt(z[pre + "ref_left"], z[pre + "ref_right"], z[pre + "cur_left"],
                                                       z[pre + "cur_right"], p_ref, p_cur, R1 @ z["R_stereo"] @ R1.T)
        want = [int(v) for v in z[pre + "oracle_inliers"]]
        if not (ginl == want and st == int(z[pre + "oracle_status"]) and np.abs(gpose - z[pre + "oracle_pose"]).max() < 1e-6):
            bad.append((pre, st, len(ginl), len(want)))
    ctx.close()
    H.diag("ransac_seeded", order=order, bad=bad)
    assert not bad, bad


@pytest.mark.parametrize("variant", ["no_nms", "topn", "binning_mask", "min_distance_8", "min_distance_3", "min_distance_1", "quality_1e-10"])
def test_detector_param_variants(variant):
    """The detector configurations of the reference's own tests (tests/testFeatureDetector.cpp:26-258:
    no NMS, TopN, Binning with a bin mask, quality 1e-10) and the uHumans2 min_distance (8, which
    reaches the 2000-corner cap) -- GPU vs oracle on a real Euroc image and a synthetic one."""
    import dataclasses
    from kimera_vio_b200.params import ANMS_TOPN
    base = FrontendParams.euroc()
    if variant == "no_nms":
        p = dataclasses.replace(base, enable_non_max_suppression=False, max_nr_keypoints_before_anms=400,
                                enable_subpixel_corner_refinement=False)
    elif variant == "topn":
        p = dataclasses.replace(base, non_max_suppression_type=ANMS_TOPN)
    elif variant == "binning_mask":
        m = np.ones((5, 4)); m[0, :] = 0; m[2, 1] = 0; m[4, 3] = 0
        p = dataclasses.replace(base, nr_horizontal_bins=4, nr_vertical_bins=5, binning_mask=m,
                                max_features_per_frame=140, quality_level=1e-10,
                                enable_subpixel_corner_refinement=False)
    elif variant == "min_distance_8":
        p = dataclasses.replace(base, min_distance=8)
    elif variant in ("min_distance_3", "min_distance_1"):
        # small cells: the greedy selection's cell grid (22 560 / 360 960 cells at 752x480) no longer fits the
        # shared-memory path and its global-scratch sections must be sized from the grid
        p = dataclasses.replace(base, min_distance=int(variant[-1]), quality_level=1e-10)
    else:
        p = dataclasses.replace(base, quality_level=1e-10)
    p, rig, ctx = H.euroc_setup(batch=1, params=p)
    g, lefts, _ = H.golden()
    s, frames = H.synth_frames(1)
    det = ofe.FeatureDetector(p)
    cam = CameraParams.euroc_left()
    for name, img in (("euroc", lefts[2]), ("synth", frames[0].left)):
        fr = ofe.Frame(0, 0, img, cam)
        need = p.max_features_per_frame
        e = det.detect_corners(fr, need)
        gq = ctx.detect(img, [], [], need)
        same_n = len(e) == len(gq)
        err = float(np.abs(gq - e).max()) if same_n and len(e) else -1.0
        H.diag("detector_variant", variant=variant, image=name, n_gpu=len(gq), n_ref=len(e), max_err=err)
        assert same_n
        assert len(e) == 0 or err <= TOL_PX
    ctx.close()
