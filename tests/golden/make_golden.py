"""Generates tests/golden/euroc_micro.npz: a few real Euroc V1_01_easy stereo pairs (the reference
ships them in tests/data/MicroEurocDataset, (c) ASL/ETHZ, see its LICENSE.md) stored PNG-compressed,
together with the ORACLE's outputs on them (cv2 4.13 in the build container).  The GPU box has no
/root/reference, so the `-m gpu` parity tests read this file; the expected outputs also pin the
oracle itself against host-dependent drift (e.g. OpenCV SIMD dispatch on a different CPU).

Run from the repo root in the build container:  python tests/golden/make_golden.py
"""
import glob
import os
import sys

import cv2
import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from kimera_vio_b200.params import CameraParams, FrontendParams  # noqa: E402
from oracle import frontend as ofe  # noqa: E402
from oracle.rig import StereoRig  # noqa: E402

SRC = "/root/reference/tests/data/MicroEurocDataset/mav0"
FRAMES = [10, 11, 12, 13, 14]


def main():
    lf = sorted(glob.glob(os.path.join(SRC, "cam0/data/*.png")))
    rf = sorted(glob.glob(os.path.join(SRC, "cam1/data/*.png")))
    out = {}
    lefts, rights, ts = [], [], []
    for k in FRAMES:
        with open(lf[k], "rb") as f:
            out["left_png_%d" % k] = np.frombuffer(f.read(), np.uint8)
        with open(rf[k], "rb") as f:
            out["right_png_%d" % k] = np.frombuffer(f.read(), np.uint8)
        lefts.append(cv2.imread(lf[k], cv2.IMREAD_GRAYSCALE))
        rights.append(cv2.imread(rf[k], cv2.IMREAD_GRAYSCALE))
        ts.append(int(os.path.basename(lf[k])[:-4]))
    out["frames"] = np.array(FRAMES)
    out["timestamps"] = np.array(ts, np.int64)
    rig = StereoRig(CameraParams.euroc_left(), CameraParams.euroc_right())
    p = FrontendParams.euroc()
    # stage outputs of the oracle on the first pair
    import zlib
    out["rect_left_crc_0"] = np.array([zlib.crc32(rig.rectify_left(lefts[0]).tobytes())], np.uint32)
    out["rect_right_crc_0"] = np.array([zlib.crc32(rig.rectify_right(rights[0]).tobytes())], np.uint32)
    out["eig_crc_0"] = np.array([zlib.crc32(cv2.cornerMinEigenVal(lefts[0], 3, ksize=3).tobytes())], np.uint32)
    det = ofe.FeatureDetector(p)
    fr = ofe.Frame(0, ts[0], lefts[0], rig.left)
    mask = det.build_mask(fr)
    raw = det.raw_feature_detection(lefts[0], mask)
    out["gftt_raw_0"] = np.array([k.pt for k in raw], np.float32)
    out["detect_0"] = det.detect_corners(fr, p.max_features_per_frame)
    # whole-sequence oracle run (identity IMU rotation is NOT given: use small synthetic rotations)
    fe = ofe.StereoFrontend(p, rig)
    seq = []
    for i, k in enumerate(FRAMES):
        R = cv2.Rodrigues(np.array([0.002 * i, -0.001 * i, 0.0015 * i]))[0] if i else np.eye(3)
        o = fe.spin(ofe.StereoFrame.make(i, ts[i], lefts[i], rights[i], rig), R)
        out["seq_R_%d" % i] = R
        out["seq_kp_%d" % i] = np.array(o.frame.left_frame.keypoints, np.float32).reshape(-1, 2)
        out["seq_lmk_%d" % i] = np.array(o.frame.left_frame.landmarks, np.int64)
        out["seq_iskf_%d" % i] = np.array([int(o.is_keyframe)])
    np.savez_compressed(os.path.join(ROOT, "tests", "golden", "euroc_micro.npz"), **out)
    print("wrote", os.path.join(ROOT, "tests", "golden", "euroc_micro.npz"))


if __name__ == "__main__":
    main()
