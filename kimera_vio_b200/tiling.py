"""Row e2 of SURVEY.md section 8(e), HOST SIDE ONLY: the band plan and the exchange protocol for one frame tiled over the
ranks of a node (C5: 3840 x 2160 over 8 GPUs).  What is here is plumbing -- which rows a rank owns, which halo rows each
stage needs, and the collective sequence of the one stage that has global dependencies, the corner detection of
cv::goodFeaturesToTrack (reference src/frontend/feature-detector/FeatureDetector.cpp:165-203):

    response rows of the band  ->  AllReduce(MAX) of the masked maximum  ->  threshold, 3 x 3 dilate test on the band
    (1-row halo of the thresholded response)  ->  gather of the candidate lists on rank 0  ->  the sequential
    min-distance selection there  ->  broadcast of the corners.

The per-band arithmetic is a call-back (`BandBackend`): tests/test_tiling_gloo.py runs the protocol under gloo with a CPU
backend and checks the corners against the whole-frame cv2.goodFeaturesToTrack, order included.  THE CUDA BACKEND IS NOT
BUILT: the kernels of libkvfe.so take whole frames (DESIGN.md section 5 explains what a banded response map needs -- the
carry of the running column sums -- and why the keypoint dimension is the better axis for LK and stereo matching).
torch.distributed is used for the collectives only (NCCL on GPUs, gloo in the test).
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import List, Optional, Protocol, Tuple

import numpy as np
import torch
import torch.distributed as dist


# ------------------------------------------------------------------------------------------------------------------
# band plan
# ------------------------------------------------------------------------------------------------------------------
def band_rows(height: int, world: int, rank: int) -> Tuple[int, int]:
    """Contiguous rows [begin, end) of `rank`; the first `height % world` bands are one row taller."""
    base, rem = divmod(height, world)
    begin = rank * base + min(rank, rem)
    return begin, begin + base + (1 if rank < rem else 0)


@dataclass(frozen=True)
class BandPlan:
    """Rows a rank owns and the halo rows each stage reads beyond them (second row of the table)."""
    rank: int
    world: int
    height: int
    begin: int
    end: int
    response_halo: int = 2      # Sobel 3 x 3 (1 row) + 3 x 3 box sum (1 row) of cv::cornerMinEigenVal
    dilate_halo: int = 1        # 3 x 3 local-maximum test on the thresholded response
    subpix_halo: int = 11       # cornerSubPix window 10 + 1 (FeatureDetector.cpp:283-296, subpixel winSize 10)
    stereo_halo: int = 5        # template rows / 2 + stripe_extra_rows (StereoMatcher.cpp:196-281, 11-row template)
    replicated_pyramid_from_level: int = 2   # LK: levels >= 2 are replicated on every rank, 0-1 exchange halos

    @staticmethod
    def make(height: int, world: int, rank: int) -> "BandPlan":
        b, e = band_rows(height, world, rank)
        return BandPlan(rank, world, height, b, e)

    def rows_with_halo(self, halo: int) -> Tuple[int, int]:
        return max(self.begin - halo, 0), min(self.end + halo, self.height)

    def response_rows(self) -> Tuple[int, int]:
        """Response rows the band must produce: its own plus the dilate halo."""
        return self.rows_with_halo(self.dilate_halo)

    def image_rows_for_response(self) -> Tuple[int, int]:
        return self.rows_with_halo(self.dilate_halo + self.response_halo)

    def lk_halo(self, level: int, win: int = 24, search: int = 8) -> int:
        """Rows of pyramid level `level` a tracked point of this band may read outside it: half a window plus the
        per-level search range, in level-0 rows ((win/2 + search) * 2^level)."""
        return (win // 2 + search) << level


def remap_source_rows(map_y: np.ndarray, begin: int, end: int, src_height: int) -> Tuple[int, int]:
    """Source-image rows the rectified rows [begin, end) read through cv::remap INTER_LINEAR: the extrema of the y map over
    the band, one more row for the bilinear tap, clamped like BORDER_REPLICATE clamps them."""
    m = map_y[begin:end]
    lo = int(np.floor(float(np.nanmin(m))))
    hi = int(np.floor(float(np.nanmax(m)))) + 2
    return max(min(lo, src_height - 1), 0), max(min(hi, src_height), 1)


# ------------------------------------------------------------------------------------------------------------------
# the detection exchange
# ------------------------------------------------------------------------------------------------------------------
class BandBackend(Protocol):
    def response_rows(self, row_begin: int, row_end: int) -> np.ndarray:
        """float32 cv::cornerMinEigenVal rows [row_begin, row_end) of the WHOLE frame's map (bit-identical to it)."""


def _all_reduce_max(value: float) -> float:
    t = torch.tensor([value], dtype=torch.float64)
    if dist.is_initialized() and dist.get_world_size() > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t[0])


def _gather_varlen(local: np.ndarray, dst: int = 0) -> Optional[List[np.ndarray]]:
    """Gather of per-rank (n_i, k) float64 arrays with different n_i: counts first, then padded blocks."""
    if not dist.is_initialized() or dist.get_world_size() == 1:
        return [local]
    world, rank = dist.get_world_size(), dist.get_rank()
    n = torch.tensor([local.shape[0]], dtype=torch.int64)
    counts = [torch.zeros(1, dtype=torch.int64) for _ in range(world)]
    dist.all_gather(counts, n)
    nmax = max(int(c[0]) for c in counts)
    pad = torch.zeros((max(nmax, 1), local.shape[1]), dtype=torch.float64)
    pad[:local.shape[0]] = torch.from_numpy(np.ascontiguousarray(local, np.float64))
    out = [torch.empty_like(pad) for _ in range(world)] if rank == dst else None
    dist.gather(pad, out, dst=dst)
    if rank != dst:
        return None
    return [o[:int(c[0])].numpy() for o, c in zip(out, counts)]


def band_candidates(resp: np.ndarray, resp_row0: int, plan: BandPlan, width: int, threshold: float,
                    mask: Optional[np.ndarray]) -> np.ndarray:
    """Steps 3-5 of cv::goodFeaturesToTrack on one band: THRESH_TOZERO at `threshold`, 3 x 3 dilate (neighbours outside
    the image do not count), keep (value, y, x) where the value is non-zero, equals the dilated value and the mask allows
    it; the frame's border rows / columns are never candidates.  resp holds rows [resp_row0, ...) incl. the dilate halo."""
    r = np.where(resp > np.float32(threshold), resp, np.float32(0.0))
    H, W = plan.height, width
    out = []
    for y in range(max(plan.begin, 1), min(plan.end, H - 1)):
        i = y - resp_row0
        row = r[i, 1:W - 1]
        nb = np.maximum.reduce([r[i + dy, 1 + dx:W - 1 + dx] for dy in (-1, 0, 1) for dx in (-1, 0, 1)])
        ok = (row != 0) & (row == nb)
        if mask is not None:
            ok &= mask[y, 1:W - 1] != 0
        xs = np.nonzero(ok)[0] + 1
        if len(xs):
            out.append(np.stack([row[xs - 1].astype(np.float64), np.full(len(xs), y, np.float64), xs.astype(np.float64)], 1))
    return np.concatenate(out) if out else np.zeros((0, 3))


def select_min_distance(cands: np.ndarray, max_corners: int, min_distance: float) -> np.ndarray:
    """The sequential part (rank 0): candidates in descending order of (value, address) -- cv's greaterThanPtr -- and a
    candidate is kept iff no already kept corner lies closer than min_distance; stops at max_corners."""
    if len(cands) == 0:
        return np.zeros((0, 2), np.float32)
    order = np.lexsort((-(cands[:, 1] * 65536 + cands[:, 2]), -cands[:, 0]))
    kept: List[Tuple[float, float]] = []
    md2 = float(min_distance) * float(min_distance)
    cell = max(int(round(min_distance)), 1)
    grid = {}
    for k in order:
        y, x = int(cands[k, 1]), int(cands[k, 2])
        good = True
        if min_distance >= 1:
            cy, cx = y // cell, x // cell
            for gy in (cy - 1, cy, cy + 1):
                for gx in (cx - 1, cx, cx + 1):
                    for (py, px) in grid.get((gy, gx), ()):
                        dy, dx = y - py, x - px
                        if dx * dx + dy * dy < md2:
                            good = False
                            break
                    if not good:
                        break
                if not good:
                    break
            if good:
                grid.setdefault((cy, cx), []).append((y, x))
        if good:
            kept.append((float(x), float(y)))
            if 0 < max_corners <= len(kept):
                break
    return np.array(kept, np.float32).reshape(-1, 2)


def tiled_good_features_to_track(backend: BandBackend, plan: BandPlan, width: int, max_corners: int, quality_level: float,
                                 min_distance: float, mask: Optional[np.ndarray] = None) -> np.ndarray:
    """cv::goodFeaturesToTrack of a frame whose rows are spread over the ranks; every rank returns the same corner list."""
    r0, r1 = plan.response_rows()
    resp = backend.response_rows(r0, r1)
    own = resp[plan.begin - r0: plan.end - r0]
    if mask is not None:
        m = mask[plan.begin:plan.end] != 0
        local_max = float(own[m].max()) if m.any() else 0.0
    else:
        local_max = float(own.max()) if own.size else 0.0
    global_max = _all_reduce_max(local_max)                                     # minMaxLoc over the whole frame
    threshold = np.float32(np.float64(global_max) * np.float64(quality_level))  # cv::threshold(eig, eig, maxVal * q, 0, TOZERO)
    # the dilate test at the first / last own row looks at the halo rows; outside the frame nothing counts
    padded = np.full((r1 - r0 + 2, width), -np.inf, np.float32)
    padded[1:-1] = resp
    cand = band_candidates(padded, r0 - 1, plan, width, float(threshold), mask)
    blocks = _gather_varlen(cand, dst=0)
    world = dist.get_world_size() if dist.is_initialized() else 1
    rank = dist.get_rank() if dist.is_initialized() else 0
    if rank == 0:
        corners = select_min_distance(np.concatenate(blocks) if blocks else np.zeros((0, 3)), max_corners, min_distance)
        n = torch.tensor([len(corners)], dtype=torch.int64)
    else:
        corners, n = None, torch.zeros(1, dtype=torch.int64)
    if world > 1:
        dist.broadcast(n, src=0)
        buf = torch.zeros((max(int(n[0]), 1), 2), dtype=torch.float32)
        if rank == 0 and len(corners):
            buf[:len(corners)] = torch.from_numpy(corners)
        dist.broadcast(buf, src=0)
        corners = buf[:int(n[0])].numpy().copy()
    return corners


# ------------------------------------------------------------------------------------------------------------------
# the other axis: one stream's keypoints spread over the ranks (every rank holds the whole frame)
# ------------------------------------------------------------------------------------------------------------------
# Pyramidal LK (Tracker.cpp:92-211) and the sparse stereo reconstruction (StereoMatcher.cpp:123-483) treat every keypoint
# on its own, so for a single large frame they shard over the keypoint list with no halo at all: a rank tracks / matches
# its block of keypoints with the ordinary stage-level entry points (kvfe_track, kvfe_sparse_stereo work on any subset of
# a frame's keypoints) and one all-gather puts the per-keypoint results back in order.  The frame itself has to be on
# every rank (an 8 MB broadcast per 4K image over NVLink).  Host logic verified under gloo (tests/test_tiling_gloo.py);
# KvfeStageBackend below wires it to libkvfe.so's verified stage calls but has never run at N > 1 on hardware.
def keypoint_block(n: int, world: int, rank: int) -> Tuple[int, int]:
    return band_rows(n, world, rank)


def _all_gather_rows(local: np.ndarray, n_total: int) -> np.ndarray:
    """Blocks of keypoint_block() order, float64 (n_i, k) -> (n_total, k) on every rank."""
    if not dist.is_initialized() or dist.get_world_size() == 1:
        return local
    world = dist.get_world_size()
    nmax = max(keypoint_block(n_total, world, r)[1] - keypoint_block(n_total, world, r)[0] for r in range(world))
    pad = torch.zeros((max(nmax, 1), local.shape[1]), dtype=torch.float64)
    pad[:local.shape[0]] = torch.from_numpy(np.ascontiguousarray(local, np.float64))
    out = [torch.empty_like(pad) for _ in range(world)]
    dist.all_gather(out, pad)
    parts = []
    for r, o in enumerate(out):
        b, e = keypoint_block(n_total, world, r)
        parts.append(o[:e - b].numpy())
    return np.concatenate(parts) if parts else np.zeros((0, local.shape[1]))


class TrackBackend(Protocol):
    def track(self, ref_xy: np.ndarray) -> Tuple[np.ndarray, np.ndarray, np.ndarray]:
        """(predicted xy, tracked xy, status u8) of the given reference keypoints (Tracker::featureTracking's LK call)."""

    def sparse_stereo(self, kps_xy: np.ndarray, versors: np.ndarray) -> dict:
        """per-keypoint arrays of StereoMatcher::sparseStereoReconstruction (left/right status, rectified xy, depth, 3-D)."""


def sharded_track(backend: TrackBackend, ref_xy: np.ndarray):
    """Every rank returns (predicted, tracked, status) for ALL keypoints; it computed only its block."""
    xy = np.asarray(ref_xy, np.float32).reshape(-1, 2)
    n = len(xy)
    world = dist.get_world_size() if dist.is_initialized() else 1
    rank = dist.get_rank() if dist.is_initialized() else 0
    b, e = keypoint_block(n, world, rank)
    pred, trk, st = backend.track(xy[b:e])
    # float32 coordinates travel as float64 (exact); one all-gather for the three results
    loc = np.concatenate([np.asarray(pred, np.float64).reshape(-1, 2), np.asarray(trk, np.float64).reshape(-1, 2),
                          np.asarray(st, np.float64).reshape(-1, 1)], 1)
    full = _all_gather_rows(loc, n)
    return full[:, 0:2].astype(np.float32), full[:, 2:4].astype(np.float32), full[:, 4].astype(np.uint8)


STEREO_FIELDS = (("left_status", 1, np.int32), ("left_rect_x", 1, np.float32), ("left_rect_y", 1, np.float32),
                 ("right_status", 1, np.int32), ("right_rect_x", 1, np.float32), ("right_rect_y", 1, np.float32),
                 ("depth", 1, np.float64), ("points_3d", 3, np.float64), ("right_x", 1, np.float32), ("right_y", 1, np.float32))


def sharded_sparse_stereo(backend: TrackBackend, kps_xy: np.ndarray, versors: np.ndarray) -> dict:
    xy = np.asarray(kps_xy, np.float32).reshape(-1, 2)
    vs = np.asarray(versors, np.float64).reshape(-1, 3)
    n = len(xy)
    world = dist.get_world_size() if dist.is_initialized() else 1
    rank = dist.get_rank() if dist.is_initialized() else 0
    b, e = keypoint_block(n, world, rank)
    res = backend.sparse_stereo(xy[b:e], vs[b:e])
    loc = np.concatenate([np.asarray(res[k], np.float64).reshape(e - b, w) for k, w, _ in STEREO_FIELDS], 1)
    full = _all_gather_rows(loc, n)
    out, c = {}, 0
    for k, w, dt in STEREO_FIELDS:
        a = full[:, c:c + w].astype(dt)
        out[k] = a.reshape(-1) if w == 1 else a
        c += w
    return out


class KvfeStageBackend:
    """The per-rank compute through libkvfe.so: a context (kimera_vio_b200.lib.Context) on this rank's GPU holding the
    whole frame's configuration; track() / sparse_stereo() are the stage-level calls on the rank's keypoint block."""

    def __init__(self, ctx, ref_img=None, cur_img=None, ref_R_cur=None, left=None, right=None):
        self.ctx, self.ref_img, self.cur_img, self.R, self.left, self.right = ctx, ref_img, cur_img, ref_R_cur, left, right

    def track(self, ref_xy):
        if len(ref_xy) == 0:
            z = np.zeros((0, 2), np.float32)
            return z, z, np.zeros(0, np.uint8)
        return self.ctx.track(self.ref_img, self.cur_img, np.eye(3) if self.R is None else self.R, ref_xy)

    def sparse_stereo(self, kps_xy, versors):
        if len(kps_xy) == 0:
            return {k: np.zeros((0, w) if w > 1 else 0, dt) for k, w, dt in STEREO_FIELDS}
        return self.ctx.sparse_stereo(self.left, self.right, kps_xy, versors)
