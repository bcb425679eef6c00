"""Row e2 (host side): the band plan and the detection exchange of kimera_vio_b200/tiling.py under gloo, world sizes 2, 3 and
4, with a CPU band backend (cv2.cornerMinEigenVal from row 0, i.e. with the carry of the running column sums): the corner list
of the tiled cv::goodFeaturesToTrack equals the whole-frame cv2.goodFeaturesToTrack, order included, with and without a mask,
for Euroc's detector parameters and for the no-minimum-distance / few-corners variants.  The CUDA band backend is not built."""
import os

import cv2
import numpy as np
import pytest
import torch.distributed as dist
import torch.multiprocessing as mp

import helpers as H
from kimera_vio_b200 import tiling as kt


class CpuBandBackend:
    """Response rows [a, b) of the whole frame's cornerMinEigenVal map: computed from row 0 so that the running column
    sums of cv::boxFilter carry the same history (a GPU rank would receive that carry from the band above)."""

    def __init__(self, img):
        self.img = img

    def response_rows(self, a, b):
        stop = min(b + 2, self.img.shape[0])
        return cv2.cornerMinEigenVal(np.ascontiguousarray(self.img[:stop]), 3, ksize=3)[a:b]


CASES = [dict(max_corners=300, quality=0.001, min_distance=20.0, masked=False),
         dict(max_corners=300, quality=0.001, min_distance=20.0, masked=True),
         dict(max_corners=40, quality=0.05, min_distance=8.0, masked=False),
         dict(max_corners=500, quality=0.001, min_distance=0.0, masked=True)]


def _mask(shape):
    m = np.full(shape, 255, np.uint8)
    rng = np.random.default_rng(3)
    for _ in range(60):
        cv2.circle(m, (int(rng.integers(0, shape[1])), int(rng.integers(0, shape[0]))), 20, 0, cv2.FILLED)
    m[200:230, :] = 0                      # a band without detections straddling rank boundaries
    return m


def _images():
    g, lefts, rights = H.golden()
    rng = np.random.default_rng(1)
    big = cv2.resize(lefts[2], (1280, 720), interpolation=cv2.INTER_CUBIC)
    big = np.clip(big.astype(np.int32) + rng.integers(-3, 4, big.shape), 0, 255).astype(np.uint8)
    return [lefts[0], rights[1], big]


def _worker(rank, world, port, q):
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    out = []
    for img in _images():
        plan = kt.BandPlan.make(img.shape[0], world, rank)
        for c in CASES:
            mask = _mask(img.shape) if c["masked"] else None
            out.append(kt.tiled_good_features_to_track(CpuBandBackend(img), plan, img.shape[1], c["max_corners"], c["quality"],
                                                       c["min_distance"], mask))
    q.put((rank, out))
    dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 3, 4])
def test_tiled_detection_equals_whole_frame(world):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29700 + (os.getpid() % 1500) + world
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    results = dict(q.get(timeout=300) for _ in range(world))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    want = []
    for img in _images():
        for c in CASES:
            mask = _mask(img.shape) if c["masked"] else None
            w = cv2.goodFeaturesToTrack(img, c["max_corners"], c["quality"], c["min_distance"], mask=mask, blockSize=3,
                                        useHarrisDetector=False)
            want.append(np.zeros((0, 2), np.float32) if w is None else w.reshape(-1, 2))
    n_nonempty = 0
    for r in range(world):
        assert len(results[r]) == len(want)
        for got, w in zip(results[r], want):
            assert got.shape == w.shape and np.array_equal(got, w)            # same corners, same order, on every rank
            n_nonempty += len(w) > 10
    assert n_nonempty >= world * 10


def test_band_plan_and_remap_halo():
    for Hh, world in ((2160, 8), (480, 3), (7, 8)):
        rows = [kt.band_rows(Hh, world, r) for r in range(world)]
        assert rows[0][0] == 0 and rows[-1][1] == Hh and all(a[1] == b[0] for a, b in zip(rows, rows[1:]))
        assert max(e - b for b, e in rows) - min(e - b for b, e in rows) <= 1
    p = kt.BandPlan.make(2160, 8, 3)
    assert (p.begin, p.end) == (810, 1080) and p.response_rows() == (809, 1081) and p.image_rows_for_response() == (807, 1083)
    assert kt.BandPlan.make(2160, 8, 0).image_rows_for_response() == (0, 273)
    assert p.lk_halo(0) == 20 and p.lk_halo(1) == 40
    # the rows of the raw image a rectified band reads: from the real Euroc maps, checked against cv2.remap on the crop
    from kimera_vio_b200.params import CameraParams
    from oracle.rig import StereoRig
    o = StereoRig(CameraParams.euroc_left(), CameraParams.euroc_right())
    g, lefts, _ = H.golden()
    full = o.rectify_left(lefts[0])
    for world, rank in ((4, 0), (4, 2), (4, 3), (8, 5)):
        b, e = kt.band_rows(o.H, world, rank)
        lo, hi = kt.remap_source_rows(o.map_ly, b, e, o.H)
        assert 0 <= lo < hi <= o.H and hi - lo < (e - b) + 60
        crop = np.ascontiguousarray(lefts[0][lo:hi])
        band = cv2.remap(crop, o.map_lx[b:e], o.map_ly[b:e] - np.float32(lo), cv2.INTER_LINEAR, borderMode=cv2.BORDER_REPLICATE)
        # rows whose taps stay inside the crop are identical; replicate-border rows at the frame edge as well
        same = (band == full[b:e]).mean()
        assert same > 0.999, (world, rank, same)


# ---- the keypoint axis: LK and sparse stereo of one frame's keypoints spread over the ranks ----------------------------
class CpuTrackBackend:
    """cv2.calcOpticalFlowPyrLK with the reference's parameters (Tracker.cpp:137-146) and the oracle's sparse stereo
    reconstruction on a block of keypoints."""

    def __init__(self, p, rig, ref_img, cur_img, left, right):
        from oracle import frontend as ofe
        self.ofe, self.p, self.rig, self.ref_img, self.cur_img, self.left, self.right = ofe, p, rig, ref_img, cur_img, left, right

    def track(self, ref_xy):
        p = self.p
        if len(ref_xy) == 0:
            z = np.zeros((0, 2), np.float32)
            return z, z, np.zeros(0, np.uint8)
        a = np.asarray(ref_xy, np.float32).reshape(-1, 1, 2)
        crit = (cv2.TERM_CRITERIA_COUNT + cv2.TERM_CRITERIA_EPS, p.klt_max_iter, p.klt_eps)
        nxt, st, _ = cv2.calcOpticalFlowPyrLK(self.ref_img, self.cur_img, a, a.copy(), winSize=(p.klt_win_size, p.klt_win_size),
                                              maxLevel=p.klt_max_level, criteria=crit, flags=cv2.OPTFLOW_USE_INITIAL_FLOW)
        return a.reshape(-1, 2), nxt.reshape(-1, 2), st.reshape(-1)

    def sparse_stereo(self, kps_xy, versors):
        ofe = self.ofe
        n = len(kps_xy)
        if n == 0:
            return {k: np.zeros((0, w) if w > 1 else 0, dt) for k, w, dt in kt.STEREO_FIELDS}
        sf = ofe.StereoFrame.make(0, 0, self.left, self.right, self.rig)
        sf.left_frame.keypoints = [(np.float32(x), np.float32(y)) for x, y in kps_xy]
        sf.left_frame.versors = [np.asarray(v, np.float64) for v in versors]
        sf.left_frame.landmarks = list(range(n))
        ofe.StereoMatcher(self.p, self.rig).sparse_stereo_reconstruction(sf)
        lr, rr = sf.left_keypoints_rectified, sf.right_keypoints_rectified
        rk = np.array(sf.right_frame.keypoints, np.float32).reshape(-1, 2)
        return dict(left_status=np.array([s for s, _ in lr], np.int32), left_rect_x=np.array([q[0] for _, q in lr], np.float32),
                    left_rect_y=np.array([q[1] for _, q in lr], np.float32), right_status=np.array([s for s, _ in rr], np.int32),
                    right_rect_x=np.array([q[0] for _, q in rr], np.float32), right_rect_y=np.array([q[1] for _, q in rr], np.float32),
                    depth=np.array(sf.keypoints_depth, np.float64), points_3d=np.array(sf.keypoints_3d, np.float64).reshape(-1, 3),
                    right_x=rk[:, 0].copy(), right_y=rk[:, 1].copy())


def _kp_setup():
    from kimera_vio_b200.params import CameraParams, FrontendParams
    from oracle import frontend as ofe
    from oracle.rig import StereoRig
    p = FrontendParams.euroc()
    rig = StereoRig(CameraParams.euroc_left(), CameraParams.euroc_right())
    g, lefts, rights = H.golden()
    kps = cv2.goodFeaturesToTrack(lefts[0], 301, 0.001, 12).reshape(-1, 2).astype(np.float32)      # 301: uneven blocks
    versors = np.array(ofe.get_bearing_vectors([tuple(q) for q in kps], rig.left, rig.R1))
    return p, rig, lefts, rights, kps, versors


def _kp_worker(rank, world, port, q):
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    p, rig, lefts, rights, kps, versors = _kp_setup()
    be = CpuTrackBackend(p, rig, lefts[0], lefts[1], lefts[0], rights[0])
    trk = kt.sharded_track(be, kps)
    ss = kt.sharded_sparse_stereo(be, kps, versors)
    empty = kt.sharded_track(be, np.zeros((0, 2), np.float32))
    few = kt.sharded_track(be, kps[:2])                     # fewer keypoints than ranks: some blocks are empty
    q.put((rank, trk, ss, len(empty[0]), few))
    dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 3])
def test_keypoint_sharded_lk_and_stereo_equal_single_rank(world):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 31300 + (os.getpid() % 1500) + world
    procs = [ctx.Process(target=_kp_worker, args=(r, world, port, q)) for r in range(world)]
    for pr in procs:
        pr.start()
    res = {}
    for _ in range(world):
        rank, trk, ss, n_empty, few = q.get(timeout=300)
        res[rank] = (trk, ss, n_empty, few)
    for pr in procs:
        pr.join(timeout=60)
        assert pr.exitcode == 0
    p, rig, lefts, rights, kps, versors = _kp_setup()
    be = CpuTrackBackend(p, rig, lefts[0], lefts[1], lefts[0], rights[0])
    w_pred, w_trk, w_st = be.track(kps)
    w_ss = be.sparse_stereo(kps, versors)
    w_few = be.track(kps[:2])
    assert w_st.sum() > 200 and (w_ss["right_status"] == 0).sum() > 150
    for r in range(world):
        (pred, trk, st), ss, n_empty, few = res[r]
        assert np.array_equal(pred, w_pred) and np.array_equal(trk, w_trk) and np.array_equal(st, w_st)
        for k, _, _ in kt.STEREO_FIELDS:
            assert np.array_equal(ss[k], w_ss[k]), k
        assert n_empty == 0
        assert np.array_equal(few[1], w_few[1]) and np.array_equal(few[2], w_few[2])
