"""Synthetic geometric scenes in the style of the reference's tests/testTracker.cpp:173-460
(AddNonPlanarInliersToFrame / AddPlanarInliersToFrame / AddOutliersToFrame and the stereo
variants).  glibc rand()/srand(3) is replaced by a seeded numpy generator, everything else
(geometry, counts, thresholds) follows the reference tests."""
import cv2
import numpy as np

from kimera_vio_b200.params import CameraParams


def expmap(w):
    return cv2.Rodrigues(np.asarray(w, np.float64).reshape(3, 1))[0]


def bearing(pt, cam: CameraParams, R=None):
    u = cv2.undistortPoints(np.array([[pt]], np.float32), cam.K, cam.D, R=R).reshape(2)
    v = np.array([float(u[0]), float(u[1]), 1.0])
    return v / np.linalg.norm(v)


def mono_scene(rng, cam: CameraParams, R, T, n_in, n_out, planar):
    """Returns (f_ref, f_cur) lists of unit bearing vectors: inliers first, then outliers."""
    Rinv, tinv = R.T, -R.T @ T
    f_ref, f_cur = [], []
    nT = np.linalg.norm(T)
    for _ in range(n_in):
        pt = (float(rng.integers(cam.width)), float(rng.integers(cam.height)))
        v = bearing(pt, cam)
        if planar:
            N = np.array([0.1, -0.1, 1.0])
            X = (nT / v.dot(N)) * v
        else:
            X = v * (nT + 9 * nT * rng.random())
        c = Rinv @ X + tinv
        f_ref.append(v)
        f_cur.append(c / np.linalg.norm(c))
    for _ in range(n_out):
        while True:
            v = bearing((float(rng.integers(cam.width)), float(rng.integers(cam.height))), cam)
            c = bearing((float(rng.integers(cam.width)), float(rng.integers(cam.height))), cam)
            proj = Rinv @ (v * nT) + tinv
            proj = proj / np.linalg.norm(proj)
            if proj.dot(c) > 0.9:
                continue
            f_ref.append(v)
            f_cur.append(c)
            break
    return np.array(f_ref), np.array(f_cur)


def stereo_scene(rng, rig, R, T, n_in, n_out, depth_range, planar=False):
    """Returns dict with ref/cur 3-D points (left-rect frame semantics as in the reference test:
    keypoints_3d hold the raw points, rectified pixel pairs come from projecting R1*p)."""
    cam = rig.left
    Rinv, tinv = R.T, -R.T @ T
    p_ref, p_cur = [], []
    for _ in range(n_in):
        v = bearing((float(rng.integers(cam.width)), float(rng.integers(cam.height))), cam)
        if planar:
            N = np.array([0.0, 0.0, 1.0])
            X = (depth_range[1] / v.dot(N)) * v
        else:
            X = v * (depth_range[0] + (depth_range[1] - depth_range[0]) * rng.random())
        p_ref.append(X)
        p_cur.append(Rinv @ X + tinv)
    for _ in range(n_out):
        while True:
            v = bearing((float(rng.integers(cam.width)), float(rng.integers(cam.height))), cam)
            c = bearing((float(rng.integers(cam.width)), float(rng.integers(cam.height))), cam)
            dr = depth_range[0] + (depth_range[1] - depth_range[0]) * rng.random()
            dc = depth_range[0] + (depth_range[1] - depth_range[0]) * rng.random()
            X, Y = v * dr, c * dc
            proj = Rinv @ X + tinv
            if np.linalg.norm(proj - Y) < 3.0 * max(0.3, 0.05 * dc):   # keep outliers clearly off-model
                continue
            p_ref.append(X)
            p_cur.append(Y)
            break
    p_ref, p_cur = np.array(p_ref), np.array(p_cur)

    def project(P):
        q = (rig.R1 @ P.T).T
        uL = rig.fx * q[:, 0] / q[:, 2] + rig.cx
        v_ = rig.fy * q[:, 1] / q[:, 2] + rig.cy
        uR = rig.fx * (q[:, 0] - rig.baseline) / q[:, 2] + rig.cx
        return (np.stack([uL, v_], 1).astype(np.float32), np.stack([uR, v_], 1).astype(np.float32))

    rl, rr = project(p_ref)
    cl, cr = project(p_cur)
    return dict(p_ref=p_ref, p_cur=p_cur, ref_left=rl, ref_right=rr, cur_left=cl, cur_right=cr)
