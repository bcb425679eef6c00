"""oracle/ransac.py -- TEST INFRASTRUCTURE (see oracle/__init__.py).

Rows a12/a13 of SURVEY.md section 8(a): geometric verification.

What is restated from the reference's own source (fully specified there):
  * 1-point stereo RANSAC ("voting")      src/frontend/Tracker.cpp:382-632
  * getPoint3AndCovariance               src/frontend/Tracker.cpp:772-818
  * Tracker::runRansac post-conditions   include/kimera-vio/frontend/Tracker.h:247-296

What is restated from a third-party dependency that is ABSENT from /root/reference and from this
image -- OpenGV (unpinned HEAD of laurentkneip/opengv, docs/kimera_vio_install.md:160), GTSAM 4.2
(Dockerfile_20_04:38) and libstdc++'s <random>:
  * opengv::sac::Ransac<P>::computeModel, SampleConsensusProblem::{getSamples,drawIndexSample}
  * relative_pose::twopt, triangulation::triangulate2, TranslationOnlySacProblem scoring
  * point_cloud::threept_arun, PointCloudSacProblem scoring
  * relative_pose::fivept_nister + CentralRelativePoseSacProblem(NISTER) disambiguation
  * gtsam::StereoCamera::backproject2 Jacobian
  * std::mt19937 + std::uniform_int_distribution<int>(0, INT_MAX) (both libstdc++ algorithms)
PARITY STATUS: the OpenGV/GTSAM half is "parity unpinned" by any reference binary; it is pinned
only by the reference's own synthetic-scene tests (tests/testTracker.cpp:704-1185: exact inlier /
outlier sets, translation within 1e-3) which tests/test_oracle_ransac.py replays, and the RNG is
pinned against this image's real libstdc++ (oracle/rng_check.cpp).

Arithmetic conventions: f64 everywhere except the f32 voting; 3-term dot products are summed as
a0*b0 + (a1*b1 + a2*b2), which is what Eigen's fixed-size redux unroller emits for size-3 vectors.
"""
from __future__ import annotations

import math
from typing import List, Optional, Sequence, Tuple

import numpy as np

# TrackingStatus -- include/kimera-vio/frontend/Tracker-definitions.h:124-130
VALID, LOW_DISPARITY, FEW_MATCHES, INVALID, DISABLED = range(5)


# ----------------------------------------------------------------------------------------------
# std::mt19937 + std::uniform_int_distribution<int>(0, INT_MAX)
# ----------------------------------------------------------------------------------------------
class StdMt19937:
    """std::mt19937 (seeded like init_genrand) -- raw 32-bit outputs."""

    def __init__(self, seed: int = 12345):
        bg = np.random.MT19937()
        bg._legacy_seeding(seed)
        self._bg = bg

    def raw(self, n: int) -> np.ndarray:
        return self._bg.random_raw(n).astype(np.uint64)


def rnd_table(n: int, seed: int = 12345, libstdcxx: str = "lemire") -> np.ndarray:
    """First n values of OpenGV's `rnd()` = uniform_int_distribution<int>(0, INT_MAX)(mt19937(seed)).

    libstdcxx = "lemire"  : GCC >= 11  (_S_nd multiply-shift; with range 2^31 this is x >> 1)
    libstdcxx = "legacy"  : GCC <  11  (rejection of raw values >= 2^31, then value / 1)
    """
    gen = StdMt19937(seed)
    if libstdcxx == "lemire":
        return (gen.raw(n) >> np.uint64(1)).astype(np.int64)
    if libstdcxx == "legacy":
        out = np.empty(0, np.int64)
        while out.size < n:
            r = gen.raw(2 * n + 64)
            out = np.concatenate([out, r[r < (1 << 31)].astype(np.int64)])
        return out[:n]
    raise ValueError(libstdcxx)


# ----------------------------------------------------------------------------------------------
# small fixed-size linear algebra in Eigen's evaluation order
# ----------------------------------------------------------------------------------------------
def dot3(a, b) -> float:
    return a[0] * b[0] + (a[1] * b[1] + a[2] * b[2])


def norm3(a) -> float:
    return math.sqrt(dot3(a, a))


def cross3(a, b) -> np.ndarray:
    return np.array([a[1] * b[2] - a[2] * b[1], a[2] * b[0] - a[0] * b[2], a[0] * b[1] - a[1] * b[0]])


def matvec3(M, v) -> np.ndarray:
    return np.array([dot3(M[0], v), dot3(M[1], v), dot3(M[2], v)])


def mattvec3(M, v) -> np.ndarray:
    return np.array([dot3(M[:, 0], v), dot3(M[:, 1], v), dot3(M[:, 2], v)])


def matmul3(A, B) -> np.ndarray:
    C = np.empty((3, 3))
    for i in range(3):
        for j in range(3):
            C[i, j] = dot3(A[i], B[:, j])
    return C


# ----------------------------------------------------------------------------------------------
# OpenGV sample-consensus machinery
# ----------------------------------------------------------------------------------------------
class SacProblem:
    sample_size = 0

    def __init__(self, n: int, rnd: np.ndarray):
        self.n = n
        self.shuffled = list(range(n))       # setUniformIndices
        self._rnd = rnd
        self._rnd_pos = 0

    def rnd(self) -> int:
        v = int(self._rnd[self._rnd_pos])
        self._rnd_pos += 1
        return v

    def draw_index_sample(self) -> List[int]:
        s, n = self.sample_size, self.n
        for i in range(s):
            j = i + (self.rnd() % (n - i))
            self.shuffled[i], self.shuffled[j] = self.shuffled[j], self.shuffled[i]
        return self.shuffled[:s]

    def get_samples(self) -> Optional[List[int]]:
        if self.n < self.sample_size:
            return None                       # iterations = INT_MAX, samples.clear()
        # max_sample_checks_ = 10, isSampleGood() is always true for these problems
        return list(self.draw_index_sample())

    def compute_model(self, sample):  # -> model or None
        raise NotImplementedError

    def scores(self, model) -> np.ndarray:
        raise NotImplementedError


def sac_ransac(problem: SacProblem, threshold: float, max_iterations: int, probability: float):
    """opengv::sac::Ransac::computeModel.  Returns (success, model, inliers, iterations)."""
    iterations = 0
    n_best = -(2 ** 31 - 1)
    k = 1.0
    skipped = 0
    max_skip = max_iterations * 10
    best_model, best_sel = None, None
    while iterations < k and skipped < max_skip:
        sel = problem.get_samples()
        if sel is None or len(sel) == 0:
            break
        model = problem.compute_model(sel)
        if model is None:
            skipped += 1
            continue
        sc = problem.scores(model)
        n_inl = int(np.count_nonzero(sc < threshold))
        if n_inl > n_best:
            n_best = n_inl
            best_model, best_sel = model, sel
            w = float(n_best) / float(problem.n)
            p_no = 1.0 - math.pow(w, float(len(sel)))
            p_no = max(np.finfo(np.float64).eps, p_no)
            p_no = min(1.0 - np.finfo(np.float64).eps, p_no)
            k = math.log(1.0 - probability) / math.log(p_no)
        iterations += 1
        if iterations > max_iterations:
            break
    if best_sel is None:
        return False, None, [], iterations
    sc = problem.scores(best_model)
    inliers = [int(i) for i in np.nonzero(sc < threshold)[0]]
    return True, best_model, inliers, iterations


def run_ransac(problem: SacProblem, threshold: float, max_iterations: int, probability: float):
    """Tracker::runRansac (Tracker.h:247-296), do_nonlinear_optimization = false."""
    ok, model, inliers, iterations = sac_ransac(problem, threshold, max_iterations, probability)
    if not ok:
        return False, np.hstack([np.eye(3), np.zeros((3, 1))]), []
    if iterations >= max_iterations and len(inliers) == 0:
        return False, np.hstack([np.eye(3), np.zeros((3, 1))]), []
    return True, model, inliers


# ----------------------------------------------------------------------------------------------
# relative pose: triangulate2 + reprojection score
# ----------------------------------------------------------------------------------------------
def triangulate2(R12, t12, f1, f2) -> np.ndarray:
    f2u = matvec3(R12, f2)
    b0, b1 = dot3(t12, f1), dot3(t12, f2u)
    a00 = dot3(f1, f1)
    a10 = dot3(f1, f2u)
    a01 = -a10
    a11 = -dot3(f2u, f2u)
    det = a00 * a11 - a10 * a01
    invdet = 1.0 / det
    i00, i10, i01, i11 = a11 * invdet, -a10 * invdet, -a01 * invdet, a00 * invdet
    l0 = i00 * b0 + i01 * b1
    l1 = i10 * b0 + i11 * b1
    xm = l0 * f1
    xn = t12 + l1 * f2u
    return (xm + xn) / 2.0


def relpose_score(R12, t12, f1, f2) -> float:
    """(1 - f1.r1) + (1 - f2.r2), r = normalised reprojections of the triangulated point."""
    X = triangulate2(R12, t12, f1, f2)
    Rt = R12.T
    tinv = -matvec3(Rt, t12)
    r1 = X
    # inverseSolution (3x4) * p_hom: Eigen's size-4 redux is (a0 + a1) + (a2 + a3)
    r2 = np.array([(Rt[k, 0] * X[0] + Rt[k, 1] * X[1]) + (Rt[k, 2] * X[2] + tinv[k] * 1.0) for k in range(3)])
    r1 = r1 / norm3(r1)
    r2 = r2 / norm3(r2)
    return (1.0 - dot3(f1, r1)) + (1.0 - dot3(f2, r2))


class Problem2d2dGivenRot(SacProblem):
    """opengv TranslationOnlySacProblem (relative_pose::twopt with unrotate=true)."""
    sample_size = 2

    def __init__(self, f_ref, f_cur, R12, rnd):
        super().__init__(len(f_ref), rnd)
        self.f1 = np.asarray(f_ref, np.float64)
        self.f2 = np.asarray(f_cur, np.float64)
        self.R12 = np.asarray(R12, np.float64)

    def compute_model(self, sample):
        i0, i1 = sample
        f1, f1p = self.f1[i0], matvec3(self.R12, self.f2[i0])
        f2, f2p = self.f1[i1], matvec3(self.R12, self.f2[i1])
        n1 = cross3(f1, f1p)
        n2 = cross3(f2, f2p)
        t = cross3(n1, n2)
        t = t / norm3(t)
        flow = f1 - f1p
        if dot3(flow, t) < 0:
            t = -t
        return np.hstack([self.R12, t.reshape(3, 1)])

    def scores(self, model):
        R, t = model[:, :3], model[:, 3]
        out = np.empty(self.n)
        for i in range(self.n):
            out[i] = relpose_score(R, t, self.f1[i], self.f2[i])
        return out


class Problem3d3d(SacProblem):
    """opengv PointCloudSacProblem (point_cloud::threept_arun)."""
    sample_size = 3

    def __init__(self, p_ref, p_cur, rnd):
        super().__init__(len(p_ref), rnd)
        self.p1 = np.asarray(p_ref, np.float64)
        self.p2 = np.asarray(p_cur, np.float64)

    def compute_model(self, sample):
        return arun(self.p1[sample], self.p2[sample])

    def scores(self, model):
        R, t = model[:, :3], model[:, 3]
        out = np.empty(self.n)
        for i in range(self.n):
            e = self.p1[i] - (matvec3(R, self.p2[i]) + t)
            out[i] = norm3(e)
        return out


def arun(p1: np.ndarray, p2: np.ndarray) -> np.ndarray:
    n = len(p1)
    c1 = np.zeros(3)
    c2 = np.zeros(3)
    for i in range(n):
        c1 = c1 + p1[i]
        c2 = c2 + p2[i]
    c1, c2 = c1 / n, c2 / n
    H = np.zeros((3, 3))
    for i in range(n):
        f = p1[i] - c1
        fp = p2[i] - c2
        H = H + np.outer(fp, f)
    U, _, Vt = np.linalg.svd(H)
    V = Vt.T
    R = V @ U.T
    if np.linalg.det(R) < 0:
        V2 = V.copy()
        V2[:, 2] = -V2[:, 2]
        R = V2 @ U.T
    t = c1 - matvec3(R, c2)
    return np.hstack([R, t.reshape(3, 1)])


# ----------------------------------------------------------------------------------------------
# 5-point Nister (opengv CentralRelativePoseSacProblem, algorithm NISTER)
# ----------------------------------------------------------------------------------------------
def _poly_mul(a, b):
    """Multiply polynomials in (x, y, z) stored as dicts {(i,j,k): coeff}."""
    out = {}
    for ka, va in a.items():
        for kb, vb in b.items():
            k = (ka[0] + kb[0], ka[1] + kb[1], ka[2] + kb[2])
            out[k] = out.get(k, 0.0) + va * vb
    return out


def _poly_add(a, b, sb=1.0):
    out = dict(a)
    for k, v in b.items():
        out[k] = out.get(k, 0.0) + sb * v
    return out


def _poly_scale(a, s):
    return {k: v * s for k, v in a.items()}


# monomial order of the 10x20 constraint matrix (degree-3 monomials in x, y, z with w = 1)
_MONO = [(3, 0, 0), (2, 1, 0), (1, 2, 0), (0, 3, 0), (2, 0, 1), (1, 1, 1), (0, 2, 1), (1, 0, 2), (0, 1, 2), (0, 0, 3),
         (2, 0, 0), (1, 1, 0), (0, 2, 0), (1, 0, 1), (0, 1, 1), (0, 0, 2), (1, 0, 0), (0, 1, 0), (0, 0, 1), (0, 0, 0)]


def fivept_essentials(f1: np.ndarray, f2: np.ndarray) -> List[np.ndarray]:
    """Essential matrices E with f1^T E f2 = 0 for 5 correspondences (Nister's constraints).

    OpenGV convention (relative_pose::fivept_nister): rows of Q are kron(f2, f1) flattened so that
    E maps frame-2 bearings into frame-1 epipolar lines:  f1^T E f2 = 0,  E = [t12]x R12.
    Solved here via the 10 cubic constraints (det E = 0, 2 E E^T E - tr(E E^T) E = 0) and the
    eigenvalues of the 10x10 action matrix; OpenGV uses a Sturm-bracketed degree-10 polynomial
    instead, so hypothesis-level parity is not claimed -- only inlier-mask parity.
    """
    Q = np.zeros((5, 9))
    for i in range(5):
        Q[i] = np.outer(f1[i], f2[i]).reshape(9)
    _, _, Vt = np.linalg.svd(Q)
    basis = [Vt[5 + k].reshape(3, 3) for k in range(4)]  # X, Y, Z, W
    # E = x X + y Y + z Z + W, entries are linear polynomials
    E = [[None] * 3 for _ in range(3)]
    for r in range(3):
        for c in range(3):
            E[r][c] = {(1, 0, 0): basis[0][r, c], (0, 1, 0): basis[1][r, c],
                       (0, 0, 1): basis[2][r, c], (0, 0, 0): basis[3][r, c]}
    # det(E)
    def det2(a, b, c, d):
        return _poly_add(_poly_mul(a, d), _poly_mul(b, c), -1.0)
    det = _poly_add(_poly_add(_poly_mul(E[0][0], det2(E[1][1], E[1][2], E[2][1], E[2][2])),
                              _poly_mul(E[0][1], det2(E[1][0], E[1][2], E[2][0], E[2][2])), -1.0),
                    _poly_mul(E[0][2], det2(E[1][0], E[1][1], E[2][0], E[2][1])))
    # EEt
    EEt = [[None] * 3 for _ in range(3)]
    for r in range(3):
        for c in range(3):
            acc = {}
            for k in range(3):
                acc = _poly_add(acc, _poly_mul(E[r][k], E[c][k]))
            EEt[r][c] = acc
    tr = _poly_add(_poly_add(EEt[0][0], EEt[1][1]), EEt[2][2])
    cons = [det]
    for r in range(3):
        for c in range(3):
            acc = {}
            for k in range(3):
                acc = _poly_add(acc, _poly_mul(EEt[r][k], E[k][c]))
            acc = _poly_add(_poly_scale(acc, 2.0), _poly_mul(tr, E[r][c]), -1.0)
            cons.append(acc)
    A = np.zeros((10, 20))
    for i, p in enumerate(cons):
        for j, m in enumerate(_MONO):
            A[i, j] = p.get(m, 0.0)
    # Gauss-Jordan on the 10 cubic monomials -> action matrix for multiplication by x
    try:
        G = np.linalg.solve(A[:, :10], A[:, 10:])
    except np.linalg.LinAlgError:
        return []
    # basis monomials B = [x^2, xy, y^2, xz, yz, z^2, x, y, z, 1] (indices 10..19 of _MONO)
    # x * B = [x^3, x^2 y, x y^2, x^2 z, x y z, x z^2, x^2, x y, x z, x]
    M = np.zeros((10, 10))
    M[0] = -G[0]   # x^3
    M[1] = -G[1]   # x^2 y
    M[2] = -G[2]   # x y^2
    M[3] = -G[4]   # x^2 z
    M[4] = -G[5]   # x y z
    M[5] = -G[7]   # x z^2
    M[6, 0] = 1.0  # x^2
    M[7, 1] = 1.0  # x y
    M[8, 3] = 1.0  # x z
    M[9, 6] = 1.0  # x
    w, V = np.linalg.eig(M)     # M B = x B: the monomial vector B is a right eigenvector
    out = []
    for k in range(10):
        if abs(w[k].imag) > 1e-9:
            continue
        v = V[:, k].real
        if abs(v[9]) < 1e-14:
            continue
        x, y, z = v[6] / v[9], v[7] / v[9], v[8] / v[9]
        out.append(x * basis[0] + y * basis[1] + z * basis[2] + basis[3])
    return out


class Problem2d2dNister(SacProblem):
    """opengv CentralRelativePoseSacProblem(NISTER): 5 points + 3 disambiguation points."""
    sample_size = 8

    def __init__(self, f_ref, f_cur, rnd):
        super().__init__(len(f_ref), rnd)
        self.f1 = np.asarray(f_ref, np.float64)
        self.f2 = np.asarray(f_cur, np.float64)

    def compute_model(self, sample):
        idx5 = list(sample[:5])
        Es = fivept_essentials(self.f1[idx5], self.f2[idx5])
        W = np.array([[0.0, -1.0, 0.0], [1.0, 0.0, 0.0], [0.0, 0.0, 1.0]])
        best_q, best = 1000000.0, None
        for E in Es:
            U, S, Vt = np.linalg.svd(E)
            scale = S[0]
            Ra = U @ W @ Vt
            Rb = U @ W.T @ Vt
            ta = scale * U[:, 2]
            tb = -ta
            if np.linalg.det(Ra) < 0:
                Ra = -Ra
            if np.linalg.det(Rb) < 0:
                Rb = -Rb
            for (R, t) in ((Ra, ta), (Ra, tb), (Rb, ta), (Rb, tb)):
                q = 0.0
                for k in sample:
                    q += relpose_score(R, t, self.f1[k], self.f2[k])
                if q < best_q:
                    best_q, best = q, np.hstack([R, t.reshape(3, 1)])
        return best

    def scores(self, model):
        R, t = model[:, :3], model[:, 3]
        out = np.empty(self.n)
        for i in range(self.n):
            out[i] = relpose_score(R, t, self.f1[i], self.f2[i])
        return out


# ----------------------------------------------------------------------------------------------
# GTSAM closed forms and the 1-point stereo voting
# ----------------------------------------------------------------------------------------------
def backproject2_jacobian(uL: float, uR: float, v: float, fx, fy, cx, cy, b) -> Tuple[np.ndarray, np.ndarray]:
    """gtsam::StereoCamera(Pose3(), K).backproject2(StereoPoint2(uL,uR,v), none, H2) (GTSAM 4.2)."""
    d = uL - uR
    z = b * fx / d
    x = z * (uL - cx) / fx
    y = z * (v - cy) / fy
    zp, xp, yp = z / d, x / d, y / d
    J = np.array([[-xp + z / fx, xp, 0.0],
                  [-yp, yp, z / fy],
                  [-zp, zp, 0.0]])
    return np.array([x, y, z]), J


def get_point3_and_covariance(left_xy, right_xy, p3d, calib, Rmat: Optional[np.ndarray]):
    """Tracker::getPoint3AndCovariance (Tracker.cpp:772-818) with stereo_point_covariance = I."""
    fx, fy, cx, cy, b = calib
    _, J = backproject2_jacobian(float(left_xy[0]), float(right_xy[0]), float(left_xy[1]), fx, fy, cx, cy, b)
    p = np.asarray(p3d, np.float64)
    if Rmat is not None:
        p = matvec3(Rmat, p)
        J = matmul3(Rmat, J)
    cov = matmul3(J, J.T)       # J * I * J^T
    return p, cov


def voting_1pt(rel_tran_f: np.ndarray, cov_f: np.ndarray, threshold: float):
    """f32 all-pairs Mahalanobis voting (Tracker.cpp:474-542). Returns (max set size, id, sets)."""
    n = len(rel_tran_f)
    th = np.float32(threshold)
    sets: List[List[int]] = [[] for _ in range(n)]
    max_size, max_id = 0, 0
    f = np.float32
    one = f(1.0)
    for i in range(n):
        sets[i].append(i)
        if i + 1 < n:
            v = rel_tran_f[i][None, :] - rel_tran_f[i + 1:]
            O = cov_f[i][None, :, :] + cov_f[i + 1:]
            v0, v1, v2 = v[:, 0], v[:, 1], v[:, 2]
            O00, O01, O02 = O[:, 0, 0], O[:, 0, 1], O[:, 0, 2]
            O10, O11, O12 = O[:, 1, 0], O[:, 1, 1], O[:, 1, 2]
            O20, O21, O22 = O[:, 2, 0], O[:, 2, 1], O[:, 2, 2]
            with np.errstate(all="ignore"):
                dinv = one / (O00 * (O11 * O22 - O12 * O21) - O10 * (O01 * O22 - O02 * O21) +
                              O20 * (O01 * O12 - O11 * O02))
                m = (dinv * v0 * (v0 * (O11 * O22 - O12 * O21) - v1 * (O01 * O22 - O02 * O21) +
                                  v2 * (O01 * O12 - O11 * O02)) +
                     dinv * v1 * (O00 * (v1 * O22 - O12 * v2) - O10 * (v0 * O22 - O02 * v2) +
                                  O20 * (v0 * O12 - v1 * O02)) +
                     dinv * v2 * (O00 * (O11 * v2 - v1 * O21) - O10 * (O01 * v2 - v0 * O21) +
                                  O20 * (O01 * v1 - O11 * v0)))
            assert m.dtype == np.float32
            for j in np.nonzero(m < th)[0]:
                jj = i + 1 + int(j)
                sets[i].append(jj)
                sets[jj].append(i)
        if len(sets[i]) > max_size:
            max_size, max_id = len(sets[i]), i
    return max_size, max_id, sets


def outlier_rejection_3d3d_given_rotation(ref_left, ref_right, cur_left, cur_right, ref_3d, cur_3d,
                                          calib, matches, R, threshold: float, min_inliers: int):
    """Tracker::geometricOutlierRejection3d3dGivenRotation (Tracker.cpp:382-632).

    ref_left/right, cur_left/right: (N,2) float32 rectified keypoints; ref_3d/cur_3d (N,3) f64;
    matches: list of (ref_idx, cur_idx).  Returns (status, pose 3x4, inliers, info 3x3).
    """
    n = len(matches)
    R = np.asarray(R, np.float64)
    rel, cov = np.zeros((n, 3)), np.zeros((n, 3, 3))
    for m, (ir, ic) in enumerate(matches):
        f_ref, c_ref = get_point3_and_covariance(ref_left[ir], ref_right[ir], ref_3d[ir], calib, None)
        f_cur, c_cur = get_point3_and_covariance(cur_left[ic], cur_right[ic], cur_3d[ic], calib, R)
        rel[m] = f_ref - f_cur
        cov[m] = c_cur + c_ref
    max_size, max_id, sets = voting_1pt(rel.astype(np.float32), cov.astype(np.float32), threshold)
    ident = np.hstack([np.eye(3), np.zeros((3, 1))])
    if max_size < 2:
        return INVALID, ident, [], np.zeros((3, 3))
    inliers = sorted(sets[max_id])
    status = VALID
    if len(inliers) < min_inliers:
        status = FEW_MATCHES
    t = np.zeros(3)
    total = np.zeros((3, 3))
    for mid in inliers:
        info = np.linalg.inv(cov[mid])
        t = t + info @ rel[mid]
        total = total + info
    t = np.linalg.inv(total) @ t
    return status, np.hstack([R, t.reshape(3, 1)]), inliers, total


def find_outliers(n_matches: int, inliers: Sequence[int]) -> List[int]:
    """Tracker::findOutliers (Tracker.cpp:836-853): set difference, ascending."""
    s = set(int(i) for i in inliers)
    return [i for i in range(n_matches) if i not in s]
