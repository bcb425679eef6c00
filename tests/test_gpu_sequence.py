"""Whole-sequence parity: the device-resident front-end FSM (kvfe_frontend_step, batch of streams)
against the oracle's StereoFrontend on the same images and IMU rotations -- every output packet
field, frame by frame (keypoints, landmark ids, ages, versors, stereo statuses, depths, 3-D points,
tracking statuses, RANSAC inlier counts, keyframe decisions)."""
import numpy as np
import pytest

import helpers as H
from kimera_vio_b200.params import CameraParams, FrontendParams
from oracle import frontend as ofe
from oracle.rig import StereoRig

pytestmark = pytest.mark.gpu
TOL_PX = 1e-3


def compare_packet(tag, pk, o, strict=True):
    """Returns a dict of mismatch counters (all zero == parity)."""
    fr = o.frame
    lf = fr.left_frame
    n = len(lf.keypoints)
    rec = dict(tag=tag, n_gpu=int(pk["n"]), n_ref=n, kf_gpu=int(pk["is_keyframe"]), kf_ref=int(o.is_keyframe))
    if pk["n"] != n or bool(pk["is_keyframe"]) != bool(o.is_keyframe):
        rec["fatal"] = 1
        return rec
    kp = np.array(lf.keypoints, np.float32).reshape(-1, 2)
    g = np.stack([pk["kp_x"], pk["kp_y"]], 1)
    rec["kp_max_err"] = float(np.abs(g - kp).max()) if n else 0.0
    rec["kp_over_tol"] = int((np.abs(g - kp).max(axis=1) > TOL_PX).sum()) if n else 0
    rec["lmk_mismatch"] = int((pk["landmark"] != np.array(lf.landmarks, np.int64)).sum())
    rec["age_mismatch"] = int((pk["age"] != np.array(lf.landmarks_age)).sum())
    rec["versor_max_err"] = float(np.abs(pk["versor"] - np.array(lf.versors).reshape(-1, 3)).max()) if n else 0.0
    rec["mono_status"] = (int(pk["mono_status"]), int(o.mono_status))
    rec["stereo_status"] = (int(pk["stereo_status"]), int(o.stereo_status))
    if o.is_keyframe:
        ls = np.array([s for s, _ in fr.left_keypoints_rectified])
        lx = np.array([q for _, q in fr.left_keypoints_rectified], np.float32).reshape(-1, 2)
        rs = np.array([s for s, _ in fr.right_keypoints_rectified])
        rx = np.array([q for _, q in fr.right_keypoints_rectified], np.float32).reshape(-1, 2)
        rec["lstat_mismatch"] = int((pk["left_status"] != ls).sum())
        rec["lrect_max_err"] = float(np.abs(np.stack([pk["left_rect_x"], pk["left_rect_y"]], 1) - lx).max())
        rec["rstat_mismatch"] = int((pk["right_status"] != rs).sum())
        rec["rrect_max_err"] = float(np.abs(np.stack([pk["right_rect_x"], pk["right_rect_y"]], 1) - rx).max())
        d = np.array(fr.keypoints_depth)
        rec["depth_max_rel"] = float((np.abs(pk["depth"] - d) / np.maximum(np.abs(d), 1e-9)).max())
        rec["p3d_max_err"] = float(np.abs(pk["point3d"] - np.array(fr.keypoints_3d).reshape(-1, 3)).max())
        rk = np.array(fr.right_frame.keypoints, np.float32).reshape(-1, 2)
        rec["right_raw_max_err"] = float(np.abs(np.stack([pk["right_x"], pk["right_y"]], 1) - rk).max())
        sm = o.smart_measurements
        rec["n_smart"] = (int(pk["n_smart"]), len(sm))
        if pk["n_smart"] == len(sm) and len(sm):
            sl = np.array([m[0] for m in sm], np.int64)
            rec["smart_lmk_mismatch"] = int((pk["smart_lmk"] != sl).sum())
            uR = np.array([m[2] for m in sm])
            rec["smart_nan_mismatch"] = int((np.isnan(pk["smart_uR"]) != np.isnan(uR)).sum())
    return rec


def packet_ok(rec):
    if rec.get("fatal"):
        return False
    ok = rec["kp_over_tol"] == 0 and rec["lmk_mismatch"] == 0 and rec["age_mismatch"] == 0
    ok &= rec["versor_max_err"] < 1e-5
    ok &= rec["mono_status"][0] == rec["mono_status"][1] and rec["stereo_status"][0] == rec["stereo_status"][1]
    if "lstat_mismatch" in rec:
        ok &= rec["lstat_mismatch"] == 0 and rec["rstat_mismatch"] == 0
        ok &= rec["lrect_max_err"] <= 2e-3 and rec["rrect_max_err"] <= TOL_PX
        ok &= rec["depth_max_rel"] < 1e-4 and rec["right_raw_max_err"] <= 2.0
        ok &= rec["n_smart"][0] == rec["n_smart"][1] and rec.get("smart_lmk_mismatch", 0) == 0
        ok &= rec.get("smart_nan_mismatch", 0) == 0
    return bool(ok)


def run_sequence(ctx, oracles, frames_per_stream, rots_per_stream, tag):
    """frames_per_stream[b][k] = (left, right, ts); rots = keyframe_R_cur provider(b, k, lkf_k)."""
    B = len(oracles)
    lkf = [0] * B
    n_frames = len(frames_per_stream[0])
    all_ok = True
    for k in range(n_frames):
        lefts = [frames_per_stream[b][k][0] for b in range(B)]
        rights = [frames_per_stream[b][k][1] for b in range(B)]
        ts = [frames_per_stream[b][k][2] for b in range(B)]
        Rs = [rots_per_stream(b, k, lkf[b]) for b in range(B)]
        pks = ctx.step(lefts, rights, ts, np.array(Rs))
        for b in range(B):
            sf = ofe.StereoFrame.make(k, ts[b], lefts[b], rights[b], oracles[b].rig)
            o = oracles[b].spin(sf, Rs[b])
            rec = compare_packet("%s/s%d/f%d" % (tag, b, k), pks[b], o)
            rec["ok"] = packet_ok(rec)
            H.diag("sequence", **rec)
            all_ok &= rec["ok"]
            if o.is_keyframe:
                lkf[b] = k
    return all_ok


def test_sequence_golden_euroc():
    """5 real Euroc pairs (tests/golden), one stream."""
    p, rig, ctx = H.euroc_setup(batch=1)
    g, lefts, rights = H.golden()
    orig = StereoRig(CameraParams.euroc_left(), CameraParams.euroc_right())
    fe = ofe.StereoFrontend(p, orig)
    ts = [int(t) for t in g["timestamps"]]
    frames = [[(lefts[i], rights[i], ts[i]) for i in range(len(lefts))]]
    ok = run_sequence(ctx, [fe], frames, lambda b, k, l: g["seq_R_%d" % k], "golden")
    ctx.close()
    assert ok


def test_sequence_synthetic_batch():
    """Two independent synthetic streams in one batch, 14 frames each (several keyframes)."""
    B, N = 2, 14
    p, rig, ctx = H.euroc_setup(batch=B)
    orig = StereoRig(CameraParams.euroc_left(), CameraParams.euroc_right())
    streams, frames = [], []
    for b in range(B):
        s, fr = H.synth_frames(N, seed=20240 + 1000 * b)
        streams.append(s)
        frames.append([(f.left, f.right, f.timestamp) for f in fr])
    oracles = [ofe.StereoFrontend(p, orig) for _ in range(B)]
    ok = run_sequence(ctx, oracles, frames, lambda b, k, l: streams[b].kf_rotation(l, k), "synth")
    ctx.close()
    assert ok


def test_sequence_d455_like_5pt_3pt():
    """The rig that switches both IMU-aided RANSAC variants off (params/D455/FrontendParams.yaml:59-60):
    5-point Nister mono + 3-point Arun stereo, one synthetic stream."""
    import dataclasses
    p = dataclasses.replace(FrontendParams.euroc(), ransac_use_2point_mono=False, ransac_use_1point_stereo=False)
    N = 10
    p, rig, ctx = H.euroc_setup(batch=1, params=p)
    orig = StereoRig(CameraParams.euroc_left(), CameraParams.euroc_right())
    s, fr = H.synth_frames(N, seed=20240)
    frames = [[(f.left, f.right, f.timestamp) for f in fr]]
    fe = ofe.StereoFrontend(p, orig)
    ok = run_sequence(ctx, [fe], frames, lambda b, k, l: s.kf_rotation(l, k), "d455")
    ctx.close()
    assert ok


def test_submit_wait_pipeline_matches_step():
    """kvfe_frontend_submit / kvfe_frontend_wait (two steps in flight, pinned and pageable output
    buffers) must produce byte-identical packets to the blocking kvfe_frontend_step."""
    import ctypes as C
    import torch
    B, N = 2, 8
    p, rig, ctxA = H.euroc_setup(batch=B)
    _, _, ctxB = H.euroc_setup(batch=B)
    streams, frames = [], []
    for b in range(B):
        s, fr = H.synth_frames(N, seed=777 + 31 * b)
        streams.append(s)
        frames.append(fr)
    # contiguous host batches (exercises the single 2-D copy upload) for A, separate arrays for B
    lkf = [0] * B
    ref, Rs_all, ts_all = [], [], []
    pkb = ctxA.packet_bytes
    for k in range(N):
        Lc = np.ascontiguousarray(np.stack([frames[b][k].left for b in range(B)]))
        Rc = np.ascontiguousarray(np.stack([frames[b][k].right for b in range(B)]))
        ts = np.array([frames[b][k].timestamp for b in range(B)], np.int64)
        Rm = np.ascontiguousarray(np.array([streams[b].kf_rotation(lkf[b], k) for b in range(B)]).reshape(B, 9))
        buf = np.empty(B * pkb, np.uint8)
        lp = (C.c_void_p * B)(*[Lc[b].ctypes.data for b in range(B)])
        rp = (C.c_void_p * B)(*[Rc[b].ctypes.data for b in range(B)])
        assert ctxA.step_raw(lp, rp, Lc.shape[2], ts, Rm, buf) == 0
        for b, pk in enumerate(ctxA.parse_packets(buf)):
            if pk["is_keyframe"]:
                lkf[b] = k
        ref.append(buf)
        Rs_all.append(Rm)
        ts_all.append(ts)
    outs = [torch.empty(B * pkb, dtype=torch.uint8).pin_memory().numpy() if k % 2 == 0 else np.empty(B * pkb, np.uint8)
            for k in range(N)]
    keep = []
    for k in range(N):
        lp = (C.c_void_p * B)(*[frames[b][k].left.ctypes.data for b in range(B)])
        rp = (C.c_void_p * B)(*[frames[b][k].right.ctypes.data for b in range(B)])
        keep.append((lp, rp))
        assert ctxB.submit_raw(lp, rp, frames[0][k].left.strides[0], ts_all[k], Rs_all[k], outs[k]) == 0
        if k >= 1:
            assert ctxB.wait() == 0
            assert np.array_equal(ctxB.packets_view(), ref[k - 1])      # zero-copy view of the same step
    assert ctxB.wait() == 0
    assert ctxB.wait() != 0                       # nothing in flight: must fail loudly
    bad = [k for k in range(N) if not np.array_equal(ref[k], outs[k])]
    H.diag("submit_wait", mismatching_frames=bad, n_kf=int(sum(l > 0 for l in lkf)))
    ctxA.close()
    ctxB.close()
    assert not bad


def test_staged_upload_matches_step():
    """kvfe_upload_frames (one H2D copy per camera for a group of contexts, issued one step ahead) +
    kvfe_frontend_submit_uploaded must give byte-identical packets to per-context kvfe_frontend_step."""
    import ctypes as C
    NCTX, N = 3, 7
    setups = [H.euroc_setup(batch=1) for _ in range(2 * NCTX)]
    ref_ctx = [s_[2] for s_ in setups[:NCTX]]
    grp_ctx = [s_[2] for s_ in setups[NCTX:]]
    pkb = ref_ctx[0].packet_bytes
    streams, frames = [], []
    for c in range(NCTX):
        s_, fr = H.synth_frames(N, seed=4242 + 17 * c)
        streams.append(s_)
        frames.append(fr)
    lib = ref_ctx[0].lib
    # reference pass (blocking step per context) -> packets and the rotations that depend on them
    lkf = [0] * NCTX
    ref, Rm_all, ts_all, Ls, Rs = [], [], [], [], []
    for k in range(N):
        L = np.ascontiguousarray(np.stack([frames[c][k].left for c in range(NCTX)]))
        R = np.ascontiguousarray(np.stack([frames[c][k].right for c in range(NCTX)]))
        ts = [np.array([frames[c][k].timestamp], np.int64) for c in range(NCTX)]
        Rm = [np.ascontiguousarray(np.asarray(streams[c].kf_rotation(lkf[c], k), np.float64).reshape(1, 9)) for c in range(NCTX)]
        row = []
        for c in range(NCTX):
            buf = np.empty(pkb, np.uint8)
            lp = (C.c_void_p * 1)(L[c].ctypes.data)
            rp = (C.c_void_p * 1)(R[c].ctypes.data)
            assert ref_ctx[c].step_raw(lp, rp, L.shape[2], ts[c], Rm[c], buf) == 0
            if ref_ctx[c].parse_packets(buf)[0]["is_keyframe"]:
                lkf[c] = k
            row.append(buf)
        ref.append(row); Rm_all.append(Rm); ts_all.append(ts); Ls.append(L); Rs.append(R)
    up = C.c_void_p()
    harr = (C.c_void_p * NCTX)(*[c.h.value for c in grp_ctx])
    assert lib.kvfe_upload_create(harr, C.c_int(NCTX), C.byref(up)) == 0
    W = Ls[0].shape[2]

    def upload(k):
        return lib.kvfe_upload_frames(up, C.c_void_p(Ls[k].ctypes.data), C.c_void_p(Rs[k].ctypes.data), C.c_size_t(W))

    bad = []
    assert upload(0) == 0
    for k in range(N):
        if k + 1 < N:
            assert upload(k + 1) == 0                       # one step ahead of the members
        if k + 2 < N:
            assert upload(k + 2) != 0                       # ring of two: must be refused
        outs = [np.empty(pkb, np.uint8) for _ in range(NCTX)]
        for c in range(NCTX):
            rc = lib.kvfe_frontend_submit_uploaded(grp_ctx[c].h, up, C.c_int(c), C.c_void_p(ts_all[k][c].ctypes.data),
                                                   C.c_void_p(Rm_all[k][c].ctypes.data), C.c_void_p(outs[c].ctypes.data))
            assert rc == 0, lib.kvfe_last_error(grp_ctx[c].h)
        for c in range(NCTX):
            assert grp_ctx[c].wait() == 0
            if not np.array_equal(ref[k][c], outs[c]):
                bad.append((k, c))
    H.diag("staged_upload", mismatches=bad)
    lib.kvfe_upload_destroy(up)
    for c in grp_ctx + ref_ctx:
        c.close()
    assert not bad


def test_submit_dev_matches_step():
    """kvfe_frontend_submit_dev (images already in device memory, two steps in flight) gives
    byte-identical packets to the blocking host-buffer step."""
    import ctypes as C
    import torch
    B, N = 2, 7
    p, rig, ctxA = H.euroc_setup(batch=B)
    _, _, ctxB = H.euroc_setup(batch=B)
    streams, frames = [], []
    for b in range(B):
        s_, fr = H.synth_frames(N, seed=5150 + 13 * b)
        streams.append(s_)
        frames.append(fr)
    pkb = ctxA.packet_bytes
    lib = ctxA.lib
    lkf = [0] * B
    ref, Rs, tss, dLs, dRs = [], [], [], [], []
    for k in range(N):
        Lc = np.ascontiguousarray(np.stack([frames[b][k].left for b in range(B)]))
        Rc = np.ascontiguousarray(np.stack([frames[b][k].right for b in range(B)]))
        ts = np.array([frames[b][k].timestamp for b in range(B)], np.int64)
        Rm = np.ascontiguousarray(np.array([streams[b].kf_rotation(lkf[b], k) for b in range(B)]).reshape(B, 9))
        buf = np.empty(B * pkb, np.uint8)
        lp = (C.c_void_p * B)(*[Lc[b].ctypes.data for b in range(B)])
        rp = (C.c_void_p * B)(*[Rc[b].ctypes.data for b in range(B)])
        assert ctxA.step_raw(lp, rp, Lc.shape[2], ts, Rm, buf) == 0
        for b, pk in enumerate(ctxA.parse_packets(buf)):
            if pk["is_keyframe"]:
                lkf[b] = k
        ref.append(buf); Rs.append(Rm); tss.append(ts)
        dLs.append(torch.from_numpy(Lc).cuda()); dRs.append(torch.from_numpy(Rc).cuda())
    torch.cuda.synchronize()
    outs = [np.empty(B * pkb, np.uint8) for _ in range(N)]
    for k in range(N):
        rc = lib.kvfe_frontend_submit_dev(ctxB.h, C.c_void_p(dLs[k].data_ptr()), C.c_void_p(dRs[k].data_ptr()),
                                          C.c_size_t(dLs[k].shape[2]), C.c_void_p(tss[k].ctypes.data),
                                          C.c_void_p(Rs[k].ctypes.data), C.c_void_p(outs[k].ctypes.data))
        assert rc == 0, lib.kvfe_last_error(ctxB.h)
        if k >= 1:
            assert ctxB.wait() == 0
    assert ctxB.wait() == 0
    bad = [k for k in range(N) if not np.array_equal(ref[k], outs[k])]
    H.diag("submit_dev", mismatching_frames=bad)
    ctxA.close()
    ctxB.close()
    assert not bad


def test_sequence_identity_rotation():
    """Euroc parameters (2-point / 1-point RANSAC enabled) with keyframe_R_cur == identity on every frame -- no
    IMU, the reference's own unit tests, exactly stationary preintegration.  Mono: outlierRejectionMono passes
    the default Pose3() and geometricOutlierRejection2d2d still picks the 2-POINT problem from the parameter
    (Tracker.cpp:248-276), now with R12 = identity -- nearly every match of a rotating camera becomes an outlier,
    which is what the reference does.  Stereo: outlierRejectionStereo really falls back to 3-point Arun
    (VisionImuFrontend.cpp:131-151).  Both behaviours must match the oracle frame by frame."""
    N = 10
    p, rig, ctx = H.euroc_setup(batch=2)
    orig = StereoRig(CameraParams.euroc_left(), CameraParams.euroc_right())
    s, fr = H.synth_frames(N, seed=20240)
    frames = [[(f.left, f.right, f.timestamp) for f in fr]] * 2
    # stream 0: identity rotations (2-point with R = I / 3-point); stream 1: the IMU rotations (2-point / 1-point), same batch
    oracles = [ofe.StereoFrontend(p, orig) for _ in range(2)]
    ok = run_sequence(ctx, oracles, frames, lambda b, k, l: np.eye(3) if b == 0 else s.kf_rotation(l, k), "identityR")
    ctx.close()
    assert ok
