"""Deterministic synthetic Euroc-shaped stereo sequences.

Scene: three textured fronto-parallel planes (Z = 1.5, 3, 6 m, nearest first, the two nearer ones
bounded in world X) observed by the distorted (radial-tangential or equidistant) stereo rig; texture = random
Gaussian blobs + band-limited noise; smooth camera motion (small rotation + translation) so that
features persist and keyframes are triggered by the reference's time / disparity logic; per-frame
sensor noise.  numpy only.  Used by bench.py (input data of the timed runs) and by the tests.
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import List, Tuple

import numpy as np

from .params import CameraParams


def _rodrigues(w: np.ndarray) -> np.ndarray:
    th = float(np.linalg.norm(w))
    if th < 1e-12:
        return np.eye(3)
    k = w / th
    K = np.array([[0, -k[2], k[1]], [k[2], 0, -k[0]], [-k[1], k[0], 0]])
    return np.eye(3) + np.sin(th) * K + (1 - np.cos(th)) * (K @ K)


def _texture(rng: np.random.Generator, size: int, n_blobs: int) -> np.ndarray:
    tex = np.zeros((size, size), np.float32)
    for _ in range(n_blobs):
        cx, cy = rng.uniform(0, size, 2)
        s = rng.uniform(2.0, 6.0)
        a = rng.uniform(-80, 80)
        r = int(np.ceil(4 * s))
        x0, x1 = max(int(cx) - r, 0), min(int(cx) + r + 1, size)
        y0, y1 = max(int(cy) - r, 0), min(int(cy) + r + 1, size)
        if x1 <= x0 or y1 <= y0:
            continue
        xs = np.arange(x0, x1, dtype=np.float32) - cx
        ys = np.arange(y0, y1, dtype=np.float32) - cy
        tex[y0:y1, x0:x1] += a * np.exp(-(ys[:, None] ** 2 + xs[None, :] ** 2) / (2 * s * s))
    # band-limited noise: white noise smoothed by a separable binomial kernel (sigma ~ 1.2)
    n = rng.standard_normal((size, size)).astype(np.float32)
    k = np.array([1, 4, 6, 4, 1], np.float32) / 16.0
    for _ in range(2):
        n = np.apply_along_axis(lambda v: np.convolve(v, k, mode="same"), 0, n)
        n = np.apply_along_axis(lambda v: np.convolve(v, k, mode="same"), 1, n)
    n = n / (n.std() + 1e-9)
    return tex + 10.0 * n + 110.0


def _undistort_grid(cam: CameraParams) -> np.ndarray:
    """Normalised undistorted coordinates (x, y) of every (distorted) pixel, shape (H, W, 2)."""
    fx, fy, cx, cy = cam.intrinsics
    k1, k2, p1, p2 = cam.distortion[:4]
    u, v = np.meshgrid(np.arange(cam.width, dtype=np.float64), np.arange(cam.height, dtype=np.float64))
    x0, y0 = (u - cx) / fx, (v - cy) / fy
    if cam.distortion_model == "equidistant":
        # theta_d = theta (1 + k1 theta^2 + k2 theta^4 + k3 theta^6 + k4 theta^8); Newton on theta, ray = tan(theta) / theta_d
        td = np.sqrt(x0 * x0 + y0 * y0)
        th = td.copy()
        for _ in range(20):
            t2 = th * th
            f = th * (1 + t2 * (k1 + t2 * (k2 + t2 * (p1 + t2 * p2)))) - td
            df = 1 + t2 * (3 * k1 + t2 * (5 * k2 + t2 * (7 * p1 + t2 * 9 * p2)))
            th = th - f / df
        sc = np.where(td > 1e-12, np.tan(th) / np.maximum(td, 1e-12), 1.0)
        return np.stack([x0 * sc, y0 * sc], -1)
    x, y = x0.copy(), y0.copy()
    for _ in range(12):
        r2 = x * x + y * y
        ic = 1.0 / (1 + (k2 * r2 + k1) * r2)
        dx = 2 * p1 * x * y + p2 * (r2 + 2 * x * x)
        dy = p1 * (r2 + 2 * y * y) + 2 * p2 * x * y
        x, y = (x0 - dx) * ic, (y0 - dy) * ic
    return np.stack([x, y], -1)


@dataclass
class SynthFrame:
    left: np.ndarray
    right: np.ndarray
    timestamp: int
    world_R_cam: np.ndarray     # left camera orientation


class SynthStream:
    """One camera stream: frame(k) renders the k-th stereo pair; kf_rotation(k_lkf, k) is the IMU-like
    relative rotation camLrectLkf_R_camLrectK expressed in the RECTIFIED left camera frame."""

    PLANES = [(1.5, -1e9, -0.45), (3.0, -0.9, 0.8), (6.0, -1e9, 1e9)]   # (Z, xmin, xmax)
    TEX_SIZE = 1536
    TEX_REL = 0.8               # texture pixels per image pixel at the plane's depth

    def __init__(self, left: CameraParams, right: CameraParams, R1: np.ndarray, seed: int = 20240,
                 rate_hz: float = 20.0, motion: float = 1.0):
        self.left, self.right, self.R1 = left, right, np.asarray(R1, np.float64)
        self.seed, self.dt_ns = seed, int(round(1e9 / rate_hz))
        self.motion = motion
        rng = np.random.default_rng(seed)
        self.tex = [_texture(rng, self.TEX_SIZE, 9000) for _ in self.PLANES]
        self.grid_l, self.grid_r = _undistort_grid(left), _undistort_grid(right)
        self.camL_T_camR = np.linalg.inv(left.T_BS) @ right.T_BS
        self.t0 = 1403715273262142976

    def pose(self, k: int) -> Tuple[np.ndarray, np.ndarray]:
        """world_R_camL, world_t_camL at frame k: slow oscillating rotation + lateral drift."""
        m = self.motion
        ang = np.deg2rad(0.4) * m * np.array([np.sin(0.11 * k), np.cos(0.07 * k), 0.5 * np.sin(0.05 * k)]) * 6.0
        R = _rodrigues(ang)
        t = m * np.array([0.02 * 12 * np.sin(0.05 * k), 0.01 * 8 * np.sin(0.031 * k + 0.5), 0.015 * 6 * np.sin(0.043 * k)])
        return R, t

    def _render(self, grid: np.ndarray, R: np.ndarray, t: np.ndarray, rng: np.random.Generator, depth_out=None) -> np.ndarray:
        H, W = grid.shape[:2]
        d = np.concatenate([grid, np.ones((H, W, 1))], -1) @ R.T      # ray directions in world
        img = np.zeros((H, W), np.float32)
        done = np.zeros((H, W), bool)
        for (Z, xmin, xmax), tex in zip(self.PLANES, self.tex):
            lam = (Z - t[2]) / d[..., 2]
            X = t[0] + lam * d[..., 0]
            Y = t[1] + lam * d[..., 1]
            ok = (~done) & (lam > 0) & (X >= xmin) & (X <= xmax)
            # texture coordinates (wrap) + bilinear sampling
            s = self.TEX_REL * self.left.intrinsics[0] / Z
            tu = (X * s + self.TEX_SIZE / 2) % (self.TEX_SIZE - 1)
            tv = (Y * s + self.TEX_SIZE / 2) % (self.TEX_SIZE - 1)
            iu, iv = np.floor(tu).astype(np.int32), np.floor(tv).astype(np.int32)
            fu, fv = (tu - iu).astype(np.float32), (tv - iv).astype(np.float32)
            iu1, iv1 = np.minimum(iu + 1, self.TEX_SIZE - 1), np.minimum(iv + 1, self.TEX_SIZE - 1)
            val = (tex[iv, iu] * (1 - fu) * (1 - fv) + tex[iv, iu1] * fu * (1 - fv) +
                   tex[iv1, iu] * (1 - fu) * fv + tex[iv1, iu1] * fu * fv)
            img = np.where(ok, val, img)
            if depth_out is not None:            # the ray is (x, y, 1) * lam in the camera frame: lam is the z-depth
                depth_out[ok] = lam[ok]
            done |= ok
        img = img + rng.normal(0.0, 1.5, img.shape).astype(np.float32)
        return np.clip(np.rint(img), 0, 255).astype(np.uint8)

    def frame(self, k: int) -> SynthFrame:
        rng = np.random.default_rng(self.seed * 1000003 + k)
        R, t = self.pose(k)
        left = self._render(self.grid_l, R, t, rng)
        Rr = R @ self.camL_T_camR[:3, :3]
        tr = t + R @ self.camL_T_camR[:3, 3]
        right = self._render(self.grid_r, Rr, tr, rng)
        return SynthFrame(left, right, self.t0 + k * self.dt_ns, R)

    def frame_with_depth(self, k: int):
        """frame(k) plus the registered metric depth image of the LEFT camera (float32, 0 where no plane is hit): the input
        of the RGB-D front-end (RgbdFrame: intensity + depth, src/frontend/RgbdFrame.cpp)."""
        rng = np.random.default_rng(self.seed * 1000003 + k)
        R, t = self.pose(k)
        depth = np.zeros(self.grid_l.shape[:2], np.float32)
        left = self._render(self.grid_l, R, t, rng, depth)
        return SynthFrame(left, left, self.t0 + k * self.dt_ns, R), depth

    def kf_rotation(self, k_lkf: int, k: int) -> np.ndarray:
        """camLrectLkf_R_camLrectK = R1 * (camL_lkf^T camL_k) * R1^T (what the IMU front-end provides)."""
        Ra, _ = self.pose(k_lkf)
        Rb, _ = self.pose(k)
        return self.R1 @ (Ra.T @ Rb) @ self.R1.T
