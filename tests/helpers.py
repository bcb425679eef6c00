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


from kimera_vio_b200.hostprobe import sobel_cpu_tail_start  # noqa: E402,F401


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


def shipped_rig(name: str):
    """FrontendParams + left/right CameraParams of one of the reference's shipped rigs (params/<name>/*.yaml),
    from tests/golden/rigs.json (values parsed from the reference YAMLs by tests/golden/make_rigs.py)."""
    with open(os.path.join(ROOT, "tests", "golden", "rigs.json")) as f:
        d = json.load(f)[name]
    fp = dict(d["frontend"])
    fp["binning_mask"] = np.asarray(fp["binning_mask"], np.float64)
    cams = []
    for side in ("left", "right"):
        c = dict(d[side])
        c["T_BS"] = np.asarray(c["T_BS"], np.float64)
        cams.append(CameraParams(**c))
    return FrontendParams(**fp), cams[0], cams[1]
