"""Builds tests/golden/tracker_scenes.npz: the synthetic scenes of the reference's RANSAC tests
(tests/testTracker.cpp:704-1185: geometricOutlierRejection2d2d / 2d2dGivenRotation / 3d3d / 3d3dGivenRotation)
regenerated from the reference's OWN random streams -- glibc srand(3) + rand() and libstdc++'s
default_random_engine + normal_distribution, produced by oracle/ref_rng.cpp with the same library calls -- and the
reference's own test camera (tests/data/ForStereoFrame/sensor{Left,Right}.yaml, 752 x 480).

`KeypointCV pt(rand() % cols, rand() % rows)` leaves the order of the two draws to the compiler: GCC evaluates
call arguments right to left (rows first), clang left to right.  Both variants are generated ("gcc", "clang");
the reference's assertions must hold on either.

The file also stores what the oracle returns on every scene (inlier lists, iteration counts, poses) at the time
of generation, so that later changes of oracle/ransac.py or of the CUDA kernels that alter a sample sequence or a
hypothesis count are caught (tests/test_oracle_ransac.py::test_reference_seeded_scenes, tests/test_gpu_stages.py).

Run from the repo root in the build container (needs /root/reference and g++):  python tests/golden/make_tracker_scenes.py
"""
import os
import subprocess
import sys

import cv2
import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from kimera_vio_b200.params import CameraParams  # noqa: E402
from oracle import ransac as rs  # noqa: E402
from oracle.rig import StereoRig  # noqa: E402

DATA = "/root/reference/tests/data/ForStereoFrame"
OUT = os.path.join(ROOT, "tests", "golden", "tracker_scenes.npz")


def build_rng():
    exe = os.path.join(ROOT, "oracle", "_build", "ref_rng")
    os.makedirs(os.path.dirname(exe), exist_ok=True)
    subprocess.check_call(["g++", "-O2", "-std=c++17", "-o", exe, os.path.join(ROOT, "oracle", "ref_rng.cpp")])
    return exe


class Rand:
    """glibc rand() after srand(3)."""
    def __init__(self, exe, n=200000):
        out = subprocess.check_output([exe, "rand", str(n)]).split()
        self.v = np.array(out[:n], np.int64)
        self.rand_max = int(out[-1])
        self.i = 0

    def __call__(self):
        self.i += 1
        return int(self.v[self.i - 1])


def normal_stream(exe, sigma, n):
    return np.array(subprocess.check_output([exe, "normal", repr(float(sigma)), str(n)]).split(), np.float64)


def expmap(w):
    return cv2.Rodrigues(np.asarray(w, np.float64).reshape(3, 1))[0]


def bearing(pt, cam):
    """UndistorterRectifier::GetBearingVector (UndistorterRectifier.cpp:62-75): undistortPoints, z = 1, normalised."""
    u = cv2.undistortPoints(np.array([[pt]], np.float32), cam.K, cam.D).reshape(2)
    v = np.array([float(u[0]), float(u[1]), 1.0])
    return v / np.linalg.norm(v)


def uncalibrate(cam, xn, yn):
    """gtsam::Cal3DS2::uncalibrate (radial-tangential k1 k2 p1 p2)."""
    k1, k2, p1, p2 = [float(c) for c in cam.D.reshape(-1)[:4]]
    fx, fy, cx, cy = cam.K[0, 0], cam.K[1, 1], cam.K[0, 2], cam.K[1, 2]
    rr = xn * xn + yn * yn
    g = 1.0 + k1 * rr + k2 * rr * rr
    dx = 2.0 * p1 * xn * yn + p2 * (rr + 2.0 * xn * xn)
    dy = 2.0 * p2 * xn * yn + p1 * (rr + 2.0 * yn * yn)
    return fx * (g * xn + dx) + cx, fy * (g * yn + dy) + cy


def keypoint(rnd, cam, order):
    if order == "gcc":          # arguments evaluated right to left: rows first
        y = rnd() % cam.height
        x = rnd() % cam.width
    else:
        x = rnd() % cam.width
        y = rnd() % cam.height
    return (float(x), float(y))


def mono_test(rnd, cam, R, T, configs, order):
    """AddNonPlanarInliersToFrame / AddPlanarInliersToFrame / AddOutliersToFrame (testTracker.cpp:212-330); the
    test's AddNoiseToFrame perturbs a COPY of every versor (`for (auto versor : ...)`) and changes nothing."""
    Rinv, tinv = R.T, -R.T @ T
    nT = float(np.linalg.norm(T))
    out = []
    for planar, n_in, n_out in configs:
        f_ref, f_cur, k_ref, k_cur = [], [], [], []
        for _ in range(n_in):
            pt = keypoint(rnd, cam, order)
            v = bearing(pt, cam)
            if planar:
                N = np.array([0.1, -0.1, 1.0])
                X = (nT / float(v.dot(N))) * v            # IntersectVersorPlane
            else:
                depth = nT + (10 * nT - nT) * (float(rnd()) / rnd.rand_max)
                X = v * depth
            c = Rinv @ X + tinv
            c = c / np.linalg.norm(c)
            f_ref.append(v); f_cur.append(c)
            k_ref.append(pt); k_cur.append(uncalibrate(cam, c[0] / c[2], c[1] / c[2]))
        for _ in range(n_out):
            while True:
                pr = keypoint(rnd, cam, order)
                pc = keypoint(rnd, cam, order)
                v, c = bearing(pr, cam), bearing(pc, cam)
                proj = Rinv @ (v * nT) + tinv
                proj = proj / np.linalg.norm(proj)
                if float(proj.dot(c)) > 0.9:
                    continue
                f_ref.append(v); f_cur.append(c); k_ref.append(pr); k_cur.append(pc)
                break
        out.append(dict(planar=planar, n_in=n_in, n_out=n_out, f_ref=np.array(f_ref), f_cur=np.array(f_cur),
                        k_ref=np.array(k_ref, np.float32), k_cur=np.array(k_cur, np.float32)))
    return out


def stereo_test(rnd, exe, rig, R, T, configs, depth_range, order):
    """AddNonPlanarInliersToStereoFrame / AddPlanarInliersToStereoFrame / AddOutliersToStereoFrame /
    AddNoiseToStereoFrame (testTracker.cpp:394-523)."""
    cam = rig.left
    Rinv, tinv = R.T, -R.T @ T
    out = []
    for planar, n_in, n_out, sigma in configs:
        p_ref, p_cur = [], []
        for _ in range(n_in):
            v = bearing(keypoint(rnd, cam, order), cam)
            if planar:
                N = np.array([0.0, 0.0, 1.0])                       # testTracker.cpp:963-965: PlaneN (0,0,1), PlaneD = depth_range[1]... see caller
                X = (depth_range[1] / float(v.dot(N))) * v
            else:
                X = v * (depth_range[0] + (depth_range[1] - depth_range[0]) * (float(rnd()) / rnd.rand_max))
            p_ref.append(X); p_cur.append(Rinv @ X + tinv)
        for _ in range(n_out):
            while True:
                v = bearing(keypoint(rnd, cam, order), cam)
                c = bearing(keypoint(rnd, cam, order), cam)
                dr = depth_range[0] + (depth_range[1] - depth_range[0]) * (float(rnd()) / rnd.rand_max)
                dc = depth_range[0] + (depth_range[1] - depth_range[0]) * (float(rnd()) / rnd.rand_max)
                X, Y = v * dr, c * dc
                proj = Rinv @ X + tinv
                if float(proj.dot(Y)) / np.linalg.norm(proj) / np.linalg.norm(Y) > 0.9:
                    continue
                p_ref.append(X); p_cur.append(Y)
                break
        p_ref, p_cur = np.array(p_ref), np.array(p_cur)
        # rectified pixel pairs come from the NOISE-FREE points (AddVersorsToStereoFrames), the noise is added to
        # keypoints_3d afterwards, each frame from a FRESH default_random_engine (same draws for ref and cur)
        q_ref, q_cur = p_ref.copy(), p_cur.copy()
        if sigma != 0:
            nz = normal_stream(exe, sigma, 3 * len(p_ref)).reshape(-1, 3)
            p_ref = p_ref + nz
            p_cur = p_cur + nz

        def project(P):
            q = (rig.R1 @ P.T).T
            uL = rig.fx * q[:, 0] / q[:, 2] + rig.cx
            v_ = rig.fy * q[:, 1] / q[:, 2] + rig.cy
            uR = rig.fx * (q[:, 0] - rig.baseline) / q[:, 2] + rig.cx
            return np.stack([uL, v_], 1).astype(np.float32), np.stack([uR, v_], 1).astype(np.float32)

        rl, rr = project(q_ref)
        cl, cr = project(q_cur)
        out.append(dict(planar=planar, n_in=n_in, n_out=n_out, sigma=sigma, p_ref=p_ref, p_cur=p_cur,
                        ref_left=rl, ref_right=rr, cur_left=cl, cur_right=cr))
    return out


def main():
    exe = build_rng()
    left = CameraParams.from_yaml(os.path.join(DATA, "sensorLeft.yaml"))
    right = CameraParams.from_yaml(os.path.join(DATA, "sensorRight.yaml"))
    rig = StereoRig(left, right)
    import dataclasses
    import json

    def enc(o):
        return o.tolist() if isinstance(o, np.ndarray) else (float(o) if isinstance(o, np.floating) else int(o))
    store = {"cams_json": np.array(json.dumps({"left": dataclasses.asdict(left), "right": dataclasses.asdict(right)}, default=enc)),
             "rig": np.array([rig.fx, rig.fy, rig.cx, rig.cy, rig.baseline]), "R1": rig.R1}
    for order in ("gcc", "clang"):
        # ---- geometricOutlierRejection2d2d (5-point Nister), testTracker.cpp:704-801
        rnd = Rand(exe)
        R, T = expmap([0.01, 0.01, 0.01]), np.array([1.0, 0, 0])
        for ci, sc in enumerate(mono_test(rnd, left, R, T, [(False, 82, 0), (False, 80, 40), (True, 80, 40)], order)):
            prob = rs.Problem2d2dNister(sc["f_ref"], sc["f_cur"], rs.rnd_table(16384))
            ok, model, inl, its = rs.sac_ransac(prob, 1e-6, 1000, 0.995)
            pre = "%s/5pt/%d/" % (order, ci)
            for k in ("f_ref", "f_cur", "k_ref", "k_cur"):
                store[pre + k] = sc[k]
            store[pre + "meta"] = np.array([sc["planar"], sc["n_in"], sc["n_out"]])
            store[pre + "oracle_inliers"] = np.array(inl, np.int32)
            store[pre + "oracle_iterations"] = np.array(its)
            store[pre + "oracle_draws"] = np.array(prob._rnd_pos)
            store[pre + "oracle_pose"] = model
            print(pre, "inliers", len(inl), "iterations", its, "draws", prob._rnd_pos)
        # ---- geometricOutlierRejection2d2dGivenRotation (2-point), testTracker.cpp:804-895
        rnd = Rand(exe)
        R, T = np.eye(3), np.array([1.0, 0, 0])
        for ci, sc in enumerate(mono_test(rnd, left, R, T, [(False, 80, 0), (False, 80, 20), (True, 80, 20)], order)):
            prob = rs.Problem2d2dGivenRot(sc["f_ref"], sc["f_cur"], R, rs.rnd_table(4096))
            ok, model, inl, its = rs.sac_ransac(prob, 1e-6, 100, 0.995)
            pre = "%s/2pt/%d/" % (order, ci)
            for k in ("f_ref", "f_cur", "k_ref", "k_cur"):
                store[pre + k] = sc[k]
            store[pre + "meta"] = np.array([sc["planar"], sc["n_in"], sc["n_out"]])
            store[pre + "oracle_inliers"] = np.array(inl, np.int32)
            store[pre + "oracle_iterations"] = np.array(its)
            store[pre + "oracle_draws"] = np.array(prob._rnd_pos)
            store[pre + "oracle_pose"] = model
            print(pre, "inliers", len(inl), "iterations", its, "draws", prob._rnd_pos)
        # ---- geometricOutlierRejection3d3d (3-point Arun) and 3d3dGivenRotation (1-point), testTracker.cpp:898-1185
        R, T = expmap([0.1, 0.1, 0.1]), np.array([rig.baseline, 0, 0])
        dr = [rig.baseline * 10, rig.baseline * 20]
        cfgs = [(False, 3, 0, 0.0), (False, 40, 0, 0.0), (False, 80, 40, 0.0), (True, 80, 40, 0.01)]
        for name in ("3pt", "1pt"):
            rnd = Rand(exe)
            for ci, sc in enumerate(stereo_test(rnd, exe, rig, R, T, cfgs, dr, order)):
                pre = "%s/%s/%d/" % (order, name, ci)
                for k in ("p_ref", "p_cur", "ref_left", "ref_right", "cur_left", "cur_right"):
                    store[pre + k] = sc[k]
                store[pre + "meta"] = np.array([sc["planar"], sc["n_in"], sc["n_out"]])
                if name == "3pt":
                    prob = rs.Problem3d3d(sc["p_ref"], sc["p_cur"], rs.rnd_table(4096))
                    ok, model, inl, its = rs.sac_ransac(prob, 0.3, 100, 0.995)
                    store[pre + "oracle_iterations"] = np.array(its)
                    store[pre + "oracle_draws"] = np.array(prob._rnd_pos)
                else:
                    calib = (rig.fx, rig.fy, rig.cx, rig.cy, rig.baseline)
                    p_ref = (rig.R1 @ sc["p_ref"].T).T
                    p_cur = (rig.R1 @ sc["p_cur"].T).T
                    Rrect = rig.R1 @ R @ rig.R1.T
                    n = len(p_ref)
                    status, model, inl, info = rs.outlier_rejection_3d3d_given_rotation(
                        sc["ref_left"], sc["ref_right"], sc["cur_left"], sc["cur_right"], p_ref, p_cur, calib,
                        [(i, i) for i in range(n)], Rrect, 1.0, 5)
                    store[pre + "oracle_status"] = np.array(status)
                store[pre + "oracle_inliers"] = np.array(inl, np.int32)
                store[pre + "oracle_pose"] = model
                print(pre, "inliers", len(inl), "of", sc["n_in"], "+", sc["n_out"])
    store["R_5pt"] = expmap([0.01, 0.01, 0.01]); store["R_stereo"] = expmap([0.1, 0.1, 0.1])
    np.savez_compressed(OUT, **store)
    print("wrote", OUT, os.path.getsize(OUT) // 1024, "KiB")


if __name__ == "__main__":
    main()
