"""Five-point relative pose oracle -- TEST INFRASTRUCTURE ONLY (never imported by the product).

Follows MotionEstimator::calcPose5PointsAlgorithm (core/visual_odometry/motion_estimator.cpp:21-123) and
findCorrectRT (:205-263).

* ``calc_pose_5point``   -- the reference's own library call, cv2.findEssentialMat(pts0, pts1, K, RANSAC, 0.999,
  thres_5p) (:41; OpenCV is third-party, cv2 4.13 here), then the SVD decomposition (:70-96) and the cheirality vote
  (:205-263) restated in numpy with the C DLT oracle (oracle/misc_oracle.c).  The RANSAC samples of OpenCV are its own:
  parity of the CUDA path with this function is statistical (inlier-set agreement, pose within a tolerance).
* ``decompose_select``   -- (:67-122) for a GIVEN essential matrix: the deterministic half, compared tightly.
* ``minimal_solutions``  -- all real essential matrices through five correspondences by the action-matrix method
  (Stewenius, Engels, Nister 2006): an algorithm different from the kernel's (Nister's elimination + degree-10
  polynomial), so agreement of the two solution sets pins the minimal solver.
* ``cv_error``           -- the error OpenCV's RANSAC thresholds (calib3d five-point.cpp, EMEstimatorCallback::computeError).
"""
import itertools

import numpy as np

from . import misc

# ---------------------------------------------------------------- polynomial algebra in (x, y, z), degree <= 3
_MONO = [e for d in (3, 2, 1, 0) for e in sorted(
    [t for t in itertools.product(range(4), repeat=3) if sum(t) == d], reverse=True)]     # graded, 20 monomials
_IDX = {m: i for i, m in enumerate(_MONO)}


def _pmul(a, b):
    out = {}
    for ma, ca in a.items():
        for mb, cb in b.items():
            m = (ma[0] + mb[0], ma[1] + mb[1], ma[2] + mb[2])
            out[m] = out.get(m, 0.0) + ca * cb
    return out


def _padd(a, b, s=1.0):
    out = dict(a)
    for m, c in b.items():
        out[m] = out.get(m, 0.0) + s * c
    return out


def _null_space(q):
    x0, y0, x1, y1 = q[:, 0], q[:, 1], q[:, 2], q[:, 3]
    Q = np.stack([x1 * x0, x1 * y0, x1, y1 * x0, y1 * y0, y1, x0, y0, np.ones(5)], 1)
    _, _, vt = np.linalg.svd(Q)
    return vt[5:9]                                                                         # 4 x 9


def minimal_solutions(q):
    """q: (5, 4) normalised (x0, y0, x1, y1) with x1^T E x0 = 0.  Returns (k, 3, 3) unit-norm real solutions."""
    q = np.asarray(q, np.float64)
    N = _null_space(q)
    var = [(1, 0, 0), (0, 1, 0), (0, 0, 1), (0, 0, 0)]
    E = [[{var[k]: N[k, 3 * i + j] for k in range(4)} for j in range(3)] for i in range(3)]
    cons = []
    det = {}
    for c in range(3):
        c1, c2 = (c + 1) % 3, (c + 2) % 3
        minor = _padd(_pmul(E[1][c1], E[2][c2]), _pmul(E[1][c2], E[2][c1]), -1.0)
        det = _padd(det, _pmul(E[0][c], minor))
    cons.append(det)
    EEt = [[{} for _ in range(3)] for _ in range(3)]
    for i in range(3):
        for j in range(3):
            for k in range(3):
                EEt[i][j] = _padd(EEt[i][j], _pmul(E[i][k], E[j][k]))
    tr = _padd(_padd(EEt[0][0], EEt[1][1]), EEt[2][2])
    for i in range(3):
        EEt[i][i] = _padd(EEt[i][i], tr, -0.5)
    for i in range(3):
        for j in range(3):
            p = {}
            for k in range(3):
                p = _padd(p, _pmul(EEt[i][k], E[k][j]))
            cons.append(p)
    A = np.zeros((10, 20))
    for r, p in enumerate(cons):
        for m, c in p.items():
            A[r, _IDX[m]] = c
    # [I | B]: every cubic monomial = -B . basis, basis = the ten monomials of degree <= 2 (columns 10..19)
    B = np.linalg.solve(A[:, :10], A[:, 10:])
    basis = _MONO[10:]
    M = np.zeros((10, 10))                                  # action of "multiply by x" on the basis
    for r, m in enumerate(basis):
        mx = (m[0] + 1, m[1], m[2])
        if sum(mx) == 3:
            M[r] = -B[_IDX[mx]]
        else:
            M[r, basis.index(mx)] = 1.0
    w, V = np.linalg.eig(M)
    ix, iy, iz, i1 = (basis.index(v) for v in var)
    sols = []
    for k in range(10):
        if abs(w[k].imag) > 1e-9 * max(1.0, abs(w[k])):
            continue
        v = V[:, k].real
        if abs(v[i1]) < 1e-14:
            continue
        x, y, z = v[ix] / v[i1], v[iy] / v[i1], v[iz] / v[i1]
        e = x * N[0] + y * N[1] + z * N[2] + N[3]
        sols.append((e / np.linalg.norm(e)).reshape(3, 3))
    return np.asarray(sols).reshape(-1, 3, 3)


def normalise(pts0, pts1, K4):
    fx, fy, cx, cy = [float(v) for v in K4]
    p0 = np.asarray(pts0, np.float64).reshape(-1, 2)
    p1 = np.asarray(pts1, np.float64).reshape(-1, 2)
    return np.stack([(p0[:, 0] - cx) / fx, (p0[:, 1] - cy) / fy, (p1[:, 0] - cx) / fx, (p1[:, 1] - cy) / fy], 1)


def cv_error(E, q):
    """(x1^T E x0)^2 / (|E x0|_xy^2 + |E^T x1|_xy^2) on normalised coordinates."""
    E = np.asarray(E, np.float64).reshape(3, 3)
    x0 = np.concatenate([q[:, 0:2], np.ones((len(q), 1))], 1)
    x1 = np.concatenate([q[:, 2:4], np.ones((len(q), 1))], 1)
    Ex0 = x0 @ E.T
    Etx1 = x1 @ E
    num = np.sum(x1 * Ex0, 1)
    return num * num / (Ex0[:, 0] ** 2 + Ex0[:, 1] ** 2 + Etx1[:, 0] ** 2 + Etx1[:, 1] ** 2)


def decompose_select(E, pts0, pts1, K4):
    """motion_estimator.cpp:67-122 for a given E10: returns (R10, t10, X0, mask_cheirality, counts[4])."""
    E = np.asarray(E, np.float32).reshape(3, 3)
    U, _, Vt = np.linalg.svd(E.astype(np.float32))
    V = Vt.T
    if np.linalg.det(U) < 0:
        U[:, 2] = -U[:, 2]
    if np.linalg.det(V) < 0:
        V[:, 2] = -V[:, 2]
    W = np.array([[0, -1, 0], [1, 0, 0], [0, 0, 1]], np.float32)
    Rs = [U @ W @ V.T, U @ W @ V.T, U @ W.T @ V.T, U @ W.T @ V.T]
    ts = [U[:, 2], -U[:, 2], U[:, 2], -U[:, 2]]
    K4 = np.asarray(K4, np.float32)
    best, counts = None, []
    mx = 0
    for R, t in zip(Rs, ts):
        X0, X1 = misc.triangulate_dlt(pts0, pts1, R.astype(np.float32), t.astype(np.float32), K4, K4)
        m = (X0[:, 2] > 0) & (X1[:, 2] > 0)
        counts.append(int(m.sum()))
        if counts[-1] > mx:
            mx = counts[-1]
            best = (R.astype(np.float32), t.astype(np.float32), X0, m)
    if best is None:
        n = len(np.asarray(pts0).reshape(-1, 2))
        return np.eye(3, dtype=np.float32), np.zeros(3, np.float32), np.zeros((n, 3), np.float32), np.ones(n, bool), counts
    return best + (counts,)


def calc_pose_5point(pts0, pts1, K4, thres_5p):
    """The reference's call sequence.  Returns (ok, R10, t10, X0, mask_inlier, E)."""
    import cv2
    p0 = np.ascontiguousarray(pts0, np.float32).reshape(-1, 2)
    p1 = np.ascontiguousarray(pts1, np.float32).reshape(-1, 2)
    fx, fy, cx, cy = [float(v) for v in K4]
    Kcv = np.array([[fx, 0, cx], [0, fy, cy], [0, 0, 1]], np.float64)
    E, inl = cv2.findEssentialMat(p0, p1, Kcv, cv2.RANSAC, 0.999, float(thres_5p))
    if E is None or E.shape != (3, 3):
        return False, None, None, None, None, None
    R, t, X0, m, _ = decompose_select(E, p0, p1, K4)
    return True, R, t, X0, m & (inl.ravel() != 0), E
