"""Stereo rectification oracle -- TEST INFRASTRUCTURE ONLY.

* rectify_maps: StereoCamera::generateStereoImagesUndistortAndRectifyMaps
  (core/visual_odometry/camera.cpp:364-546), float32 in the reference's operation order.  Eigen's 3x3 products /
  inverse are restated as plain k-ordered sums / cofactors (third-party, unpinned); everything else is the cited code.
* remap_linear: what StereoCamera::rectifyStereoImages + the convertTo(CV_8UC1) of StereoVO do to an 8-bit image
  (camera.cpp:300-336, stereo_vo.cpp:416-421): cv::remap(CV_32FC1 image, CV_32FC1 maps, INTER_LINEAR, BORDER_CONSTANT 0)
  followed by saturate_cast<uchar>.  cv::remap is third-party OpenCV (pinned here against cv2 4.13 by
  tests/test_oracle_rectify.py): map coordinates are rounded to 1/32 px (cvRound(x * 32), round half to even), the four
  taps are weighted with the float products (1-fy)(1-fx), (1-fy)fx, fy(1-fx), fy fx and summed left to right.
"""
import numpy as np

from . import mono_step as omono

f32 = np.float32


def _mul3v(M, v):
    M = np.asarray(M, f32)
    return [f32(f32(f32(M[r, 0] * v[0]) + f32(M[r, 1] * v[1])) + f32(M[r, 2] * v[2])) for r in range(3)]


def rectify_maps(K_l, D_l, K_r, D_r, T_lr, w, h):
    """Returns dict(map_lu, map_lv, map_ru, map_rv [h,w] f32, K_rect [4] f32, T_lr_rect [4,4] f32)."""
    K_l, K_r = np.asarray(K_l, f32), np.asarray(K_r, f32)
    D_l, D_r = np.asarray(D_l, f32), np.asarray(D_r, f32)          # k1 k2 p1 p2 k3 (camera.cpp:30-35)
    T_lr = np.asarray(T_lr, f32)
    R_0r, t_0r = T_lr[:3, :3], T_lr[:3, 3]
    R_l0 = np.eye(3, dtype=f32)
    R_r0 = R_0r.T.copy()
    k_l = np.array([0, 0, 1], f32)
    k_r = R_0r[:, 2]
    k_n = ((k_l + k_r) * f32(0.5)).astype(f32)

    def unit(v):
        n = f32(np.sqrt(f32(f32(f32(v[0] * v[0]) + f32(v[1] * v[1])) + f32(v[2] * v[2]))))
        return (v / n).astype(f32)

    def cross(a, b):
        return np.array([f32(f32(a[1] * b[2]) - f32(a[2] * b[1])), f32(f32(a[2] * b[0]) - f32(a[0] * b[2])),
                         f32(f32(a[0] * b[1]) - f32(a[1] * b[0]))], f32)
    k_n = unit(k_n)
    i_n = unit(t_0r.astype(f32))
    j_n = unit(cross(k_n, i_n))
    k_n = unit(cross(i_n, j_n))
    R_0n = np.stack([i_n, j_n, k_n], 1).astype(f32)
    f_n = f32(f32(K_l[0] + K_r[0]) * f32(0.5))
    centu, centv = f32(f32(w) * f32(0.5)), f32(f32(h) * f32(0.5))
    K_rect = np.array([[f_n, 0, centu], [0, f_n, centv], [0, 0, 1]], f32)
    K_rect_inv = omono.inv3_f32(K_rect)
    M = omono.mul3_f32(R_0n, K_rect_inv)                               # R_0n * K_rect_inv, then * p_n
    us = np.arange(w, dtype=f32) + f32(1.0)
    vs = np.arange(h, dtype=f32) + f32(1.0)
    U, V = np.meshgrid(us, vs)
    one = f32(1.0)
    P0 = [((M[r, 0] * U + M[r, 1] * V) + M[r, 2] * one).astype(f32) for r in range(3)]
    out = {}
    for name, Rc, K, D in (("l", R_l0, K_l, D_l), ("r", R_r0, K_r, D_r)):
        X = [((Rc[r, 0] * P0[0] + Rc[r, 1] * P0[1]) + Rc[r, 2] * P0[2]).astype(f32) for r in range(3)]
        with np.errstate(divide="ignore", invalid="ignore"):
            x, y = (X[0] / X[2]).astype(f32), (X[1] / X[2]).astype(f32)
        k1, k2, p1, p2, k3 = D[0], D[1], D[2], D[3], D[4]
        xx, yy = x * x, y * y
        xy2 = x * y * f32(2.0)
        r2 = xx + yy
        r4 = r2 * r2
        r6 = r4 * r2
        r_radial = ((one + k1 * r2) + k2 * r4) + k3 * r6
        x_dist = (x * r_radial + p1 * xy2) + p2 * (r2 + f32(2.0) * xx)
        y_dist = (y * r_radial + p2 * xy2) + p1 * (r2 + f32(2.0) * yy)
        out["map_" + name + "u"] = ((x_dist * K[0] + K[2]) - one).astype(f32)
        out["map_" + name + "v"] = ((y_dist * K[1] + K[3]) - one).astype(f32)
    R_ln = omono.mul3_f32(R_l0, R_0n)
    t_rect = _mul3v(R_ln.T.copy(), t_0r)
    T = np.eye(4, dtype=f32)
    T[:3, 3] = t_rect
    out["K_rect"] = np.array([f_n, f_n, centu, centv], f32)
    out["T_lr_rect"] = T
    return out


def remap_linear(img, map_u, map_v):
    """cv::remap(float(img), map_u, map_v, INTER_LINEAR, BORDER_CONSTANT 0) -> convertTo(CV_8UC1). img: u8 [h, w]."""
    img = np.asarray(img)
    h, w = img.shape
    src = img.astype(f32)
    mu, mv = np.asarray(map_u, f32), np.asarray(map_v, f32)
    with np.errstate(invalid="ignore", over="ignore"):
        sx = np.rint(mu * f32(32.0))
        sy = np.rint(mv * f32(32.0))
    # cvRound of a non-finite / huge float is INT_MIN on x86 (cvtss2si); such pixels end up outside the image
    bad = ~np.isfinite(sx) | ~np.isfinite(sy) | (np.abs(sx) >= 2.0 ** 31) | (np.abs(sy) >= 2.0 ** 31)
    sx = np.where(bad, -2.0 ** 31, sx).astype(np.int64)
    sy = np.where(bad, -2.0 ** 31, sy).astype(np.int64)
    fx, fy = (sx & 31).astype(f32) * f32(1.0 / 32.0), (sy & 31).astype(f32) * f32(1.0 / 32.0)
    ix = np.clip(sx >> 5, -32768, 32767)                                # saturate_cast<short>
    iy = np.clip(sy >> 5, -32768, 32767)
    one = f32(1.0)
    w00, w01 = (one - fy) * (one - fx), (one - fy) * fx
    w10, w11 = fy * (one - fx), fy * fx

    def tap(yy, xx):
        ok = (xx >= 0) & (xx < w) & (yy >= 0) & (yy < h)
        return np.where(ok, src[np.clip(yy, 0, h - 1), np.clip(xx, 0, w - 1)], f32(0)).astype(f32)
    val = ((tap(iy, ix) * w00 + tap(iy, ix + 1) * w01) + tap(iy + 1, ix) * w10) + tap(iy + 1, ix + 1) * w11
    return np.clip(np.rint(val.astype(f32)), 0, 255).astype(np.uint8)


def undistort_maps(K4, D5, w, h):
    """Camera::generateImageUndistortMaps (core/visual_odometry/camera.cpp:57-87): float variables, double literals (the
    promotions of `2.0 * x * y`, `1.0 + k1 * r2 + ...`, `2.0 * xx` restated)."""
    f32, f64 = np.float32, np.float64
    fx, fy, cx, cy = [f32(v) for v in K4]
    k1, k2, p1, p2, k3 = [f32(v) for v in D5]
    fxinv, fyinv = f32(f32(1.0) / fx), f32(f32(1.0) / fy)
    u = np.arange(w, dtype=f32)[None, :].repeat(h, 0)
    v = np.arange(h, dtype=f32)[:, None].repeat(w, 1)
    y = ((v - cy).astype(f32) * fyinv).astype(f32)
    x = ((u - cx).astype(f32) * fxinv).astype(f32)
    xy2 = ((f64(2.0) * x.astype(f64)) * y.astype(f64)).astype(f32)
    xx, yy = (x * x).astype(f32), (y * y).astype(f32)
    r2 = (xx + yy).astype(f32)
    r4 = (r2 * r2).astype(f32)
    r6 = (r4 * r2).astype(f32)
    r_radial = (((f64(1.0) + (k1 * r2).astype(f32).astype(f64)) + (k2 * r4).astype(f32).astype(f64)) + (k3 * r6).astype(f32).astype(f64)).astype(f32)
    x_dist = (((x * r_radial).astype(f32) + (p1 * xy2).astype(f32)).astype(f32).astype(f64) +
              f64(p2) * (r2.astype(f64) + f64(2.0) * xx.astype(f64))).astype(f32)
    y_dist = (((y * r_radial).astype(f32).astype(f64) + f64(p1) * (r2.astype(f64) + f64(2.0) * yy.astype(f64))) +
              (p2 * xy2).astype(f32).astype(f64)).astype(f32)
    return (cx + (x_dist * fx).astype(f32)).astype(f32), (cy + (y_dist * fy).astype(f32)).astype(f32)
