"""oracle/mesher.py -- TEST INFRASTRUCTURE ONLY.  CPU restatement of the 2-D Delaunay mesh the
Mesher builds right downstream of the front-end packet (config C3 names it).
No CUDA counterpart exists yet: this is the oracle the next round's kernel will be checked against.

Mesher::createMesh2dImpl   src/mesh/Mesher.cpp:1712-1817  (cv::Subdiv2D incremental Delaunay)
Mesher::createMesh2D       src/mesh/Mesher.cpp:1819-1845
Mesher::createMesh2dStereo src/mesh/Mesher.cpp:1849-1886
Pinned by tests/testMesher.cpp:147-197 (tests/test_oracle_pins.py).
"""
from typing import List, Sequence, Tuple

import cv2
import numpy as np

KP_VALID = 0


def _rect_contains(w: int, h: int, x: float, y: float) -> bool:
    # cv::Rect2f(0, 0, w, h).contains: x <= pt.x < x + width, same for y
    return 0.0 <= x < float(w) and 0.0 <= y < float(h)


def create_mesh_2d_impl(img_size: Tuple[int, int], keypoints: Sequence[Tuple[float, float]]) -> np.ndarray:
    """img_size = (width, height). Returns (T, 6) float32: x0 y0 x1 y1 x2 y2 per triangle."""
    w, h = img_size
    if len(keypoints) == 0:                                            # :1716
        return np.zeros((0, 6), np.float32)
    rect = (0, 0, int(w), int(h))
    subdiv = cv2.Subdiv2D(rect)                                        # :1722-1723
    inside = []
    for x, y in keypoints:                                             # :1733-1749
        x, y = float(np.float32(x)), float(np.float32(y))
        if _rect_contains(w, h, x, y) and x >= 0.0 and y >= 0.0:
            inside.append((x, y))
    if inside:
        subdiv.insert(inside)                                          # :1753
    tri = subdiv.getTriangleList()                                     # :1774
    if tri is None or len(tri) == 0:
        return np.zeros((0, 6), np.float32)
    tri = np.asarray(tri, np.float32).reshape(-1, 6)
    good = [t for t in tri                                             # :1786-1790
            if _rect_contains(w, h, t[0], t[1]) and _rect_contains(w, h, t[2], t[3]) and _rect_contains(w, h, t[4], t[5])]
    return np.asarray(good, np.float32).reshape(-1, 6)


def create_mesh_2d(img_size: Tuple[int, int], keypoints, landmarks, selected_indices) -> np.ndarray:
    """Mesher::createMesh2D(const Frame&, selected_indices)."""
    assert len(keypoints) == len(landmarks)
    w, h = img_size
    sel = []
    for i in selected_indices:
        x, y = keypoints[i]
        # cv::Rect2i::contains(Point2f): integer rectangle, float point
        if landmarks[i] != -1 and 0 <= x < w and 0 <= y < h:
            sel.append((x, y))
    return create_mesh_2d_impl(img_size, sel)


def create_mesh_2d_stereo(img_size, landmarks, right_status, keypoints, points_3d=None):
    """Mesher::createMesh2dStereo: keypoints with a VALID right match and a live landmark; also returns
    the (landmark id, 3-D point in the left camera) pairs when points_3d is given."""
    assert len(landmarks) == len(right_status)
    kps, lmk3d = [], []
    for i in range(len(landmarks)):
        if right_status[i] == KP_VALID and landmarks[i] != -1:
            kps.append(keypoints[i])
            if points_3d is not None:
                lmk3d.append((int(landmarks[i]), np.asarray(points_3d[i], np.float64)))
    return create_mesh_2d_impl(img_size, kps), lmk3d
