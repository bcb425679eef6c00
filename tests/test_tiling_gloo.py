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
