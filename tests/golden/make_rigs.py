"""Generates tests/golden/rigs.json: the parameter VALUES of the reference's shipped rigs -- parsed from
params/<rig>/{FrontendParams,LeftCameraParams,RightCameraParams}.yaml with kimera_vio_b200.params
(the same keys the reference parses) -- so that the GPU box, which has no /root/reference, can run the
front-end with the real uHumans2 / D455 / uHumans1 / Euroc configurations instead of scaled Euroc values.

Run from the repo root in the build container:  python tests/golden/make_rigs.py
"""
import dataclasses
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from kimera_vio_b200.params import CameraParams, FrontendParams  # noqa: E402

REF = "/root/reference/params"
RIGS = ["Euroc", "uHumans2", "uHumans1", "D455", "RealSenseIR"]


def enc(o):
    if isinstance(o, np.ndarray):
        return o.tolist()
    if isinstance(o, (np.integer,)):
        return int(o)
    if isinstance(o, (np.floating,)):
        return float(o)
    raise TypeError(type(o))


def main():
    out = {}
    for r in RIGS:
        d = os.path.join(REF, r)
        out[r] = {"frontend": dataclasses.asdict(FrontendParams.from_yaml(os.path.join(d, "FrontendParams.yaml"))),
                  "left": dataclasses.asdict(CameraParams.from_yaml(os.path.join(d, "LeftCameraParams.yaml"))),
                  "right": dataclasses.asdict(CameraParams.from_yaml(os.path.join(d, "RightCameraParams.yaml")))}
    with open(os.path.join(ROOT, "tests", "golden", "rigs.json"), "w") as f:
        json.dump(out, f, default=enc, indent=1)
    print("wrote rigs.json:", {r: (out[r]["left"]["width"], out[r]["left"]["height"], out[r]["left"]["distortion_model"]) for r in RIGS})


if __name__ == "__main__":
    main()
