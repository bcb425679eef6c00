"""Shared helpers for the GPU parity tests."""
import json
import os

import cv2
import numpy as np

from kimera_vio_b200 import lib as kl
from kimera_vio_b200.params import CameraParams, FrontendParams
from kimera_vio_b200.rig import StereoRigSetup
from kimera_vio_b200.synth import SynthStream

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OUT = os.path.join(ROOT, "gpurun_out")
GOLDEN = os.path.join(ROOT, "tests", "golden", "euroc_micro.npz")


def diag(name, **kw):
    """Append a diagnostics record to gpurun_out/diag.jsonl (comes back from the GPU box)."""
    os.makedirs(OUT, exist_ok=True)

    def conv(v):
        if isinstance(v, np.ndarray):
            return v.tolist()
        if isinstance(v, (np.integer,)):
            return int(v)
        if isinstance(v, (np.floating,)):
            return float(v)
        return v
    with open(os.path.join(OUT, "diag.jsonl"), "a") as f:
        f.write(json.dumps({"test": name, **{k: conv(v) for k, v in kw.items()}}) + "\n")


def sobel_cpu_tail_start(width: int) -> int:
    """Detects where THIS host's cv2 switches to the scalar (non-FMA) Sobel row-filter tail
     by probing cv2.Sobel on a random image; -1 when no tail is used."""
    rng = np.random.default_rng(5)
    img = rng.integers(0, 256, (64, width), dtype=np.uint8)
    dy = cv2.Sobel(img, cv2.CV_32F, 0, 1, ksize=3, scale=1 / 3060.0)
    I = img.astype(np.float32)
    s = np.float32(1 / 3060.0)
    s2 = np.float32(2) * s
    xs = np.arange(width)
    xm, xp = np.abs(xs - 1), np.where(xs + 1 >= width, 2 * (width - 1) - (xs + 1), xs + 1)
    ys = np.arange(64)
    ym, yp = np.abs(ys - 1), np.where(ys + 1 >= 64, 2 * 63 - (ys + 1), ys + 1)
    t0 = s * I[:, xm]
    t1 = (I.astype(np.float64) * np.float64(s2) + t0.astype(np.float64)).astype(np.float32)
    t_fma = (I[:, xp].astype(np.float64) * np.float64(s) + t1.astype(np.float64)).astype(np.float32)
    t_nofma = (s * I[:, xm] + s2 * I) + s * I[:, xp]
    dy_fma = t_fma[yp] - t_fma[ym]
    dy_no = t_nofma[yp] - t_nofma[ym]
    col_fma_ok = np.all(dy_fma == dy, axis=0)
    col_no_ok = np.all(dy_no == dy, axis=0)
    # the scalar tail is the suffix of columns that only the non-FMA formula explains
    if col_fma_ok.all():
        return -1
    start = int(np.argmin(col_fma_ok))
    assert col_no_ok[start:].all() and col_fma_ok[:start].all(), "unexpected cv2.Sobel arithmetic on this host"
    return start


def euroc_setup(batch=1, params=None, **cfg_kw):
    p = params or FrontendParams.euroc()
    rig = StereoRigSetup(CameraParams.euroc_left(), CameraParams.euroc_right())
    tail = sobel_cpu_tail_start(rig.W)
    cfg = kl.make_config(p, rig.W, rig.H, batch=batch, sobel_cpu_tail_start=tail, **cfg_kw)
    ctx = kl.Context(cfg, rig.to_c())
    return p, rig, ctx


def golden():
    g = np.load(GOLDEN)
    frames = [int(k) for k in g["frames"]]
    lefts = [cv2.imdecode(g["left_png_%d" % k], cv2.IMREAD_GRAYSCALE) for k in frames]
    rights = [cv2.imdecode(g["right_png_%d" % k], cv2.IMREAD_GRAYSCALE) for k in frames]
    return g, lefts, rights


_synth_cache = {}


def synth_frames(n, seed=20240):
    key = (n, seed)
    if key not in _synth_cache:
        rig = StereoRigSetup(CameraParams.euroc_left(), CameraParams.euroc_right())
        s = SynthStream(CameraParams.euroc_left(), CameraParams.euroc_right(), rig.R1, seed=seed)
        _synth_cache[key] = (s, [s.frame(k) for k in range(n)])
    return _synth_cache[key]
