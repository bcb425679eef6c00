"""include/kvfe_shim.hpp with DATA: tests/native/shim_run.cpp (C++, built here with g++ -Werror) drives every method of
the shim on a real stereo pair; the same calls through ctypes must give the same bytes."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

import helpers as H
from kimera_vio_b200 import build as kb
from kimera_vio_b200 import lib as kl

pytestmark = pytest.mark.gpu
ROOT = H.ROOT


def read_blobs(path):
    d, raw = {}, open(path, "rb").read()
    i = 0
    while i < len(raw):
        nl = int.from_bytes(raw[i:i + 4], "little"); i += 4
        name = raw[i:i + nl].decode(); i += nl
        es = int.from_bytes(raw[i:i + 4], "little"); i += 4
        n = int.from_bytes(raw[i:i + 8], "little"); i += 8
        d[name] = (es, raw[i:i + es * n]); i += es * n
    return d


def test_shim_methods_with_data(tmp_path):
    p, rig, ctx = H.euroc_setup(batch=1)
    g, lefts, rights = H.golden()
    L, R, L2 = lefts[0], rights[0], lefts[1]
    cfg = kl.make_config(p, rig.W, rig.H, batch=1, sobel_cpu_tail_start=H.sobel_cpu_tail_start(rig.W))
    inp, outp, exe = tmp_path / "in.bin", tmp_path / "out.bin", tmp_path / "shim_run"
    with open(inp, "wb") as f:
        f.write(bytes(cfg)); f.write(bytes(rig.to_c())); f.write(L.tobytes()); f.write(R.tobytes()); f.write(L2.tobytes())
    so = kb.build()
    libdir = os.path.dirname(so)
    r = subprocess.run(["g++", "-std=c++17", "-O1", "-Wall", "-Wextra", "-Werror", "-I", os.path.join(ROOT, "include"), "-o", str(exe),
                        os.path.join(ROOT, "tests", "native", "shim_run.cpp"), "-L", libdir, "-lkvfe", "-Wl,-rpath," + libdir],
                       capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    run = subprocess.run([str(exe), str(inp), str(outp)], capture_output=True, text=True)
    assert run.returncode == 0, run.stdout + run.stderr
    b = read_blobs(outp)

    def arr(name, dt):
        es, raw = b[name]
        a = np.frombuffer(raw, dt)
        assert a.itemsize == es, name
        return a

    def xy(name):
        return np.stack([arr(name + ".x", np.float32), arr(name + ".y", np.float32)], 1)
    # the same sequence through ctypes
    Lr, Rr = ctx.rectify_pair(L, R)
    assert np.array_equal(arr("rect_left", np.uint8).reshape(L.shape), Lr) and np.array_equal(arr("rect_right", np.uint8).reshape(L.shape), Rr)
    kps = ctx.detect(L, need=150)
    assert np.array_equal(xy("detect"), kps) and len(kps) > 100
    mask = np.full(L.shape, 255, np.uint8); mask[:, rig.W // 3:rig.W // 2] = 0
    assert np.array_equal(xy("detect_masked"), ctx.detect_masked(L, mask, need=150))
    und = ctx.undistort_keypoints(0, True, True, kps)
    assert np.array_equal(xy("undistort"), und)
    vers = ctx.bearing_vectors(kps)
    assert np.array_equal(arr("versors", np.float64).reshape(-1, 3), vers)
    lst, lrect = ctx.undistort_rectify_left_keypoints(kps)
    assert np.array_equal(arr("left_status", np.int32), lst) and np.array_equal(xy("left_rect"), lrect)
    cst, ckp = ctx.check_rectified_keypoints(0, kps, und, 1.0)
    assert np.array_equal(arr("check_status", np.int32), cst) and np.array_equal(xy("check_kps"), ckp)
    pred, trk, tst = ctx.track(L, L2, np.eye(3), kps)
    assert np.array_equal(xy("track"), trk) and np.array_equal(arr("track_status", np.uint8), tst)
    ss = ctx.sparse_stereo(L, R, kps, vers)
    assert np.array_equal(arr("ss_left_status", np.int32), ss["left_status"]) and np.array_equal(arr("ss_right_status", np.int32), ss["right_status"])
    assert np.array_equal(arr("ss_depth", np.float64), ss["depth"]) and np.array_equal(arr("ss_points", np.float64).reshape(-1, 3), ss["points_3d"])
    rst, rrect = ctx.right_keypoints_rectified(Lr, Rr, lst, lrect)
    assert np.array_equal(arr("right_status", np.int32), rst) and np.array_equal(xy("right_rect"), rrect)
    rst2, depth = ctx.depth_from_rectified_matches(lst, lrect[:, 0], rst, rrect[:, 0])
    assert np.array_equal(arr("depth", np.float64), depth) and np.array_equal(arr("right_status_after_depth", np.int32), rst2)
    assert np.array_equal(xy("right_unrect"), ctx.distort_unrectify_keypoints(1, rst2, rrect))
    m = np.nonzero(tst)[0].astype(np.int32)
    ok, med = ctx.compute_median_disparity(kps, trk, np.stack([m, m], 1))
    assert np.array_equal(arr("median", np.float64), np.array([1.0 if ok else 0.0, med]))
    lre = np.stack([ss["left_rect_x"], ss["left_rect_y"]], 1)
    rre = np.stack([ss["right_rect_x"], ss["right_rect_y"]], 1)
    p3, cov = ctx.point3_and_covariance(lre, rre, ss["points_3d"], np.eye(3))
    assert np.array_equal(arr("p3", np.float64).reshape(-1, 3), p3, equal_nan=True) and np.array_equal(arr("cov", np.float64).reshape(-1, 3, 3), cov, equal_nan=True)
    v2 = ctx.bearing_vectors(trk)
    st2, _, inl2 = ctx.ransac_mono(vers[m], v2[m], np.eye(3))
    assert int(arr("ransac2_status", np.int32)[0]) == st2 and list(arr("ransac2_inliers", np.int32)) == inl2
    # the RGB-D additions (RgbdFrame in the shim)
    depth16 = (500 + 20 * L.astype(np.int32)).astype(np.uint16)
    dpar = kl.DepthParams(0, 0.1, 0.001, 2.0, 4.0)
    dmask = ctx.depth_detection_mask(depth16, dpar)
    assert np.array_equal(arr("depth_mask", np.uint8).reshape(L.shape), dmask) and 0 < dmask.mean() < 255
    frs, frx, fdep, fp3, frk = ctx.rgbd_fill_stereo_frame(depth16, dpar, kps, lst, lrect, vers)
    assert np.array_equal(arr("fill_right_status", np.int32), frs) and np.array_equal(xy("fill_right_rect"), frx)
    assert np.array_equal(arr("fill_depth", np.float64), fdep) and np.array_equal(arr("fill_points", np.float64).reshape(-1, 3), fp3)
    assert np.array_equal(xy("fill_right_kps"), frk) and (frs == 0).sum() > 20 and (frs == 3).sum() > 0
    # host bookkeeping added to the Tracker shim: findMatchingKeypoints / findMatchingStereoKeypoints
    n_k = len(kps)
    want = [i for i in range(n_k) if i % 3 != 0]
    assert list(arr("find_matches_ref", np.int32)) == want
    wants = [i for i in want if ss["right_status"][i] == 0]
    assert list(arr("find_stereo_matches_ref", np.int32)) == wants and list(arr("find_stereo_matches_cur", np.int32)) == wants
    ctx.close()
    for name in ("ransac5_status", "ransac3_inliers", "ransac1_inliers", "outliers", "lmk_ref_after", "matches_after", "ss_right_kps.x"):
        assert name in b
