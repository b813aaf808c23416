"""Rectification on the device (camera.cpp:300-546): maps against the float32 restatement, remap bit-exact against the
cv2-pinned oracle, and the rectified image feeding the pyramid path."""
import numpy as np
import pytest

from oracle import klt as oklt
from oracle import rectify as orect
from visual_odometry_ros_b200 import capi, synth

pytestmark = pytest.mark.gpu


def _rig():
    K_l = np.array([458.654, 457.296, 367.215, 248.375], np.float32)
    K_r = np.array([457.587, 456.134, 379.999, 255.238], np.float32)
    D_l = np.array([-0.28340811, 0.07395907, 0.00019359, 1.76187114e-05, 0.0], np.float32)
    D_r = np.array([-0.28368365, 0.07451284, -0.00010473, -3.55590700e-05, 0.0], np.float32)
    T = np.eye(4)
    T[:3, :3] = synth.so3_exp([0.004, -0.012, 0.002])
    T[:3, 3] = [0.11, 0.0006, -0.0009]
    return K_l, D_l, K_r, D_r, T.astype(np.float32), 752, 480


def test_rectify_maps_and_remap():
    K_l, D_l, K_r, D_r, T, w, h = _rig()
    rng = np.random.default_rng(8)
    ctx = capi.Context(device=0, max_w=w, max_h=h, n_slots=2, max_feat=256)
    with pytest.raises(capi.VoError):
        ctx.upload_image_rectified(0, 0, np.zeros((h, w), np.uint8))            # maps not built yet
    K_rect, T_rect = ctx.rectify_init(K_l, D_l, K_r, D_r, T, w, h)
    o = orect.rectify_maps(K_l, D_l, K_r, D_r, T, w, h)
    assert np.array_equal(K_rect, o["K_rect"])
    assert np.abs(T_rect - o["T_lr_rect"]).max() <= 1e-7
    for right, (nu, nv) in enumerate((("map_lu", "map_lv"), ("map_ru", "map_rv"))):
        mu, mv = ctx.read_rectify_maps(right)
        du, dv = np.abs(mu - o[nu]).max(), np.abs(mv - o[nv]).max()
        print(f"maps cam {right}: max |du| {du:.2e} |dv| {dv:.2e}, identical {np.mean(mu == o[nu]):.4f}")
        assert du <= 1e-4 and dv <= 1e-4                                          # same FP32 operation order (host 3x3 algebra may differ by an ulp)
        img = synth.textured_image(rng, w, h)
        ctx.upload_image_rectified(right, right, img)
        ctx.build_pyramids(np.array([right], np.int32), 3, True)
        got, _ = ctx.read_pyramid_level(right, 0)
        ref = orect.remap_linear(img, mu, mv)                                      # oracle remap through the DEVICE maps
        assert np.array_equal(got, ref), int((got != ref).sum())
        # and the pyramid built from the rectified image is the cv2 pyramid of that image
        lv, dvv = oklt.build_pyramid(ref, 21, 2)
        g1, d1 = ctx.read_pyramid_level(right, 1)
        assert np.array_equal(g1, lv[1]) and np.array_equal(d1, dvv[1])
    with pytest.raises(capi.VoError):
        ctx.upload_image_rectified(0, 0, np.zeros((h - 2, w), np.uint8))          # camera.cpp:308 size check
    ctx.close()


def test_mono_undistortion_maps_and_image(gpu_ctx):
    """Camera::generateImageUndistortMaps + undistortImage (camera.cpp:57-87, 163-183) for MonoVO's flagDoUndistortion:
    the map is bit-identical to the numpy restatement (float / double promotions included), the undistorted image bit-exact
    with the cv2-pinned remap oracle."""
    from oracle import rectify as orect
    w, h = 752, 480
    K4 = np.array([458.654, 457.296, 367.215, 248.375], np.float32)
    D5 = np.array([-0.28340811, 0.07395907, 0.00019359, 1.76187114e-05, 0.0], np.float32)
    gpu_ctx_local = capi.Context(device=0, max_w=w, max_h=h, n_slots=2, max_feat=1024)
    gpu_ctx_local.undistort_init(K4, D5, w, h)
    mu, mv = gpu_ctx_local.read_rectify_maps(0)
    ou, ov = orect.undistort_maps(K4, D5, w, h)
    assert np.array_equal(mu, ou) and np.array_equal(mv, ov)
    img = synth.textured_image(np.random.default_rng(5), w, h)
    gpu_ctx_local.upload_image_rectified(0, 0, img)
    gpu_ctx_local.build_pyramids(np.array([0], np.int32), 1, True)
    got = gpu_ctx_local.read_pyramid_level(0, 0)[0]
    assert np.array_equal(got, orect.remap_linear(img, ou, ov))
    assert np.abs(got.astype(int) - img.astype(int)).mean() > 1.0          # the distortion does move pixels
    gpu_ctx_local.close()
