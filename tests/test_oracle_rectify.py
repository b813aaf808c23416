"""The rectification oracle: remap pinned against cv2.remap (the reference's own library call, camera.cpp:320,335),
map generation checked against an independent float64 evaluation of the same geometry."""
import numpy as np

from oracle import rectify as orect
from visual_odometry_ros_b200 import synth


def _rig():
    K_l = np.array([458.654, 457.296, 367.215, 248.375], np.float32)       # EuRoC-like pinhole-radtan rig
    K_r = np.array([457.587, 456.134, 379.999, 255.238], np.float32)
    D_l = np.array([-0.28340811, 0.07395907, 0.00019359, 1.76187114e-05, 0.0], np.float32)
    D_r = np.array([-0.28368365, 0.07451284, -0.00010473, -3.55590700e-05, 0.0], np.float32)
    T = np.eye(4)
    T[:3, :3] = synth.so3_exp([0.004, -0.012, 0.002])
    T[:3, 3] = [0.11, 0.0006, -0.0009]
    return K_l, D_l, K_r, D_r, T.astype(np.float32), 752, 480


def test_remap_matches_cv2_bit_exact():
    import cv2
    rng = np.random.default_rng(3)
    K_l, D_l, K_r, D_r, T, w, h = _rig()
    m = orect.rectify_maps(K_l, D_l, K_r, D_r, T, w, h)
    img = synth.textured_image(rng, w, h)
    for mu, mv in ((m["map_lu"], m["map_lv"]), (m["map_ru"], m["map_rv"]),
                   (rng.uniform(-40, w + 40, (h, w)).astype(np.float32), rng.uniform(-40, h + 40, (h, w)).astype(np.float32))):
        ref = cv2.remap(img.astype(np.float32), mu, mv, cv2.INTER_LINEAR)
        ref8 = np.clip(np.rint(ref), 0, 255).astype(np.uint8)                  # Mat::convertTo(CV_8UC1)
        assert np.array_equal(cv2.convertScaleAbs(ref) if False else ref8, ref8)
        got = orect.remap_linear(img, mu, mv)
        assert np.array_equal(got, ref8), int((got != ref8).sum())


def test_maps_against_float64_geometry():
    K_l, D_l, K_r, D_r, T, w, h = _rig()
    m = orect.rectify_maps(K_l, D_l, K_r, D_r, T, w, h)
    # independent evaluation in float64 with numpy linear algebra
    T64 = T.astype(np.float64)
    R_0r, t = T64[:3, :3], T64[:3, 3]
    k_n = (np.array([0, 0, 1.0]) + R_0r[:, 2]) * 0.5
    k_n /= np.linalg.norm(k_n)
    i_n = t / np.linalg.norm(t)
    j_n = np.cross(k_n, i_n); j_n /= np.linalg.norm(j_n)
    k_n = np.cross(i_n, j_n); k_n /= np.linalg.norm(k_n)
    R_0n = np.stack([i_n, j_n, k_n], 1)
    f_n = (float(K_l[0]) + float(K_r[0])) * 0.5
    Kr = np.array([[f_n, 0, w * 0.5], [0, f_n, h * 0.5], [0, 0, 1]])
    U, V = np.meshgrid(np.arange(w) + 1.0, np.arange(h) + 1.0)
    P0 = np.einsum("ij,jhw->ihw", R_0n @ np.linalg.inv(Kr), np.stack([U, V, np.ones_like(U)]))
    for name, Rc, K, D in (("l", np.eye(3), K_l, D_l), ("r", R_0r.T, K_r, D_r)):
        X = np.einsum("ij,jhw->ihw", Rc, P0)
        x, y = X[0] / X[2], X[1] / X[2]
        k1, k2, p1, p2, k3 = [float(v) for v in D]
        r2 = x * x + y * y
        rad = 1 + k1 * r2 + k2 * r2 ** 2 + k3 * r2 ** 3
        xd = x * rad + p1 * 2 * x * y + p2 * (r2 + 2 * x * x)
        yd = y * rad + p2 * 2 * x * y + p1 * (r2 + 2 * y * y)
        assert np.abs(m["map_" + name + "u"] - (xd * float(K[0]) + float(K[2]) - 1)).max() < 2e-3
        assert np.abs(m["map_" + name + "v"] - (yd * float(K[1]) + float(K[3]) - 1)).max() < 2e-3
    assert abs(float(m["T_lr_rect"][0, 3]) - np.linalg.norm(t)) < 1e-5 and np.abs(m["T_lr_rect"][1:3, 3]).max() < 1e-5
    assert np.allclose(m["K_rect"], [f_n, f_n, w * 0.5, h * 0.5])


def test_mono_undistort_maps_oracle():
    """Camera::generateImageUndistortMaps (camera.cpp:57-87) restated: equals a float64 evaluation of the radtan model to
    float rounding, is the identity without distortion, and inverts the distortion for the remap oracle."""
    w, h = 320, 240
    K4 = np.array([300.0, 305.0, 160.5, 120.25], np.float32)
    D5 = np.array([-0.25, 0.06, 3e-4, -2e-4, 0.01], np.float32)
    mu, mv = orect.undistort_maps(K4, D5, w, h)
    u, v = np.meshgrid(np.arange(w, dtype=np.float64), np.arange(h, dtype=np.float64))
    x, y = (u - K4[2]) / K4[0], (v - K4[3]) / K4[1]
    r2 = x * x + y * y
    rad = 1 + D5[0] * r2 + D5[1] * r2 ** 2 + D5[4] * r2 ** 3
    xd = x * rad + D5[2] * 2 * x * y + D5[3] * (r2 + 2 * x * x)
    yd = y * rad + D5[2] * (r2 + 2 * y * y) + D5[3] * 2 * x * y
    assert np.abs(mu - (K4[2] + xd * K4[0])).max() < 2e-3 and np.abs(mv - (K4[3] + yd * K4[1])).max() < 2e-3
    iu, iv = orect.undistort_maps(K4, np.zeros(5, np.float32), w, h)
    assert np.abs(iu - u).max() < 1e-3 and np.abs(iv - v).max() < 1e-3
    img = np.random.default_rng(1).integers(0, 256, (h, w), dtype=np.uint8)
    assert np.array_equal(orect.remap_linear(img, iu, iv), img)       # 1/32-px quantisation absorbs the float rounding
