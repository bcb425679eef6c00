"""Generates tests/golden/rgbd_pair.npz from the reference's RGB-D test data (tests/data/ForRgbd: depth_img_0.tiff,
CV_32FC1 720x480, left_img_0.png) and its camera file sensorLeft.yaml, so that the GPU box -- which has no
/root/reference -- can run the RGB-D stage tests on the real pair.  The depth image is stored losslessly (float32).

Run from the repo root in the build container:  python tests/golden/make_rgbd.py
"""
import dataclasses
import json
import os
import sys

import cv2
import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from kimera_vio_b200.params import CameraParams  # noqa: E402

REF = "/root/reference/tests/data/ForRgbd"


def main():
    depth = cv2.imread(os.path.join(REF, "depth_img_0.tiff"), cv2.IMREAD_UNCHANGED)
    assert depth.dtype == np.float32 and depth.shape == (480, 720)
    # UtilsOpenCV::ReadAndConvertToGrayScale (UtilsOpenCV.cpp:390-403): imread, cvtColor BGR2GRAY for 3 channels
    img = cv2.imread(os.path.join(REF, "left_img_0.png"), cv2.IMREAD_ANYCOLOR)
    if img.ndim == 3:
        img = cv2.cvtColor(img, cv2.COLOR_BGR2GRAY)
    cam = CameraParams.from_yaml(os.path.join(REF, "sensorLeft.yaml"))
    d = dataclasses.asdict(cam)
    d["T_BS"] = cam.T_BS.tolist()
    np.savez_compressed(os.path.join(ROOT, "tests", "golden", "rgbd_pair.npz"), depth=depth, left=img, camera=json.dumps(d))
    print("wrote rgbd_pair.npz", depth.shape, img.shape, os.path.getsize(os.path.join(ROOT, "tests", "golden", "rgbd_pair.npz")))


if __name__ == "__main__":
    main()
