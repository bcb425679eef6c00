"""Builds tests/golden/_euroc95.npz: ALL 95 stereo pairs of the reference's tests/data/MicroEurocDataset
(Euroc V1_01_easy, (c) ASL/ETHZ, see its LICENSE.md), PNG bytes as shipped, plus the frame-to-frame
rotation camLrectKm1_R_camLrectK integrated from the dataset's own gyroscope samples (imu0/data.csv,
no bias correction -- an input like any other: the oracle and the GPU path receive the same matrices).

The file is ~70 MB and therefore NOT committed (.gitignore); it is rebuilt by __graft_entry__.build()
whenever /root/reference is present and travels to the GPU box with the repo snapshot.  The long-sequence
parity test (tests/test_gpu_long.py) skips cleanly when it is absent.

Run from the repo root in the build container:  python tests/golden/make_euroc95.py
"""
import glob
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from kimera_vio_b200.params import CameraParams  # noqa: E402
from kimera_vio_b200.rig import StereoRigSetup  # noqa: E402

SRC = "/root/reference/tests/data/MicroEurocDataset/mav0"
OUT = os.path.join(ROOT, "tests", "golden", "_euroc95.npz")


def so3_exp(w):
    th = float(np.linalg.norm(w))
    if th < 1e-12:
        return np.eye(3)
    k = w / th
    K = np.array([[0, -k[2], k[1]], [k[2], 0, -k[0]], [-k[1], k[0], 0]])
    return np.eye(3) + np.sin(th) * K + (1 - np.cos(th)) * (K @ K)


def main():
    lf = sorted(glob.glob(os.path.join(SRC, "cam0/data/*.png")))
    rf = sorted(glob.glob(os.path.join(SRC, "cam1/data/*.png")))
    assert len(lf) == len(rf) == 95
    ts = np.array([int(os.path.basename(f)[:-4]) for f in lf], np.int64)
    imu = np.loadtxt(os.path.join(SRC, "imu0/data.csv"), delimiter=",", skiprows=1)
    it, gyro = imu[:, 0].astype(np.int64), imu[:, 1:4]
    left = CameraParams.euroc_left()
    rig = StereoRigSetup(left, CameraParams.euroc_right())
    body_R_cam = left.T_BS[:3, :3]
    out = {"timestamps": ts}
    rel = [np.eye(3)]
    for k in range(1, len(ts)):
        sel = np.nonzero((it >= ts[k - 1]) & (it < ts[k]))[0]
        dR = np.eye(3)
        for i in sel:                          # midpoint-free forward integration of the raw gyro samples
            dt = (min(it[i + 1], ts[k]) - it[i]) * 1e-9
            dR = dR @ so3_exp(gyro[i] * dt)
        rel.append(rig.R1 @ (body_R_cam.T @ dR @ body_R_cam) @ rig.R1.T)
    out["rel_R"] = np.stack(rel)
    for k in range(len(lf)):
        with open(lf[k], "rb") as f:
            out["left_png_%d" % k] = np.frombuffer(f.read(), np.uint8)
        with open(rf[k], "rb") as f:
            out["right_png_%d" % k] = np.frombuffer(f.read(), np.uint8)
    np.savez(OUT, **out)
    print("wrote", OUT, os.path.getsize(OUT) >> 20, "MiB")


if __name__ == "__main__":
    main()
