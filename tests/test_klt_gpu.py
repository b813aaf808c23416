"""GPU parity of K-pyr / K-klt against the oracle (cv2 4.13 = the reference's own library call,
feature_tracker.cpp:29..186, and the C restatement oracle/klt_oracle.c), through the C ABI."""
import numpy as np
import pytest

from visual_odometry_ros_b200 import synth

pytestmark = pytest.mark.gpu

TOL_PX = 0.01          # BASELINE.json north_star: tracked positions within 0.01 px
STATUS_AGREE = 0.999   # >= 99.9 % agreement on track-status flags


@pytest.fixture(scope="module")
def case():
    return synth.klt_stereo_case(seed=2002, n=2000)


def test_pyramid_bit_exact(gpu_ctx, case):
    import cv2
    gpu_ctx.upload_image(0, case["left"])
    gpu_ctx.build_pyramids([0], 5, True)
    nl, pyr = cv2.buildOpticalFlowPyramid(case["left"], (21, 21), 4, withDerivatives=True)
    for l in range(5):
        img, der = gpu_ctx.read_pyramid_level(0, l)
        assert np.array_equal(img, pyr[2 * l]), f"level {l} image differs"
        assert np.array_equal(der, pyr[2 * l + 1]), f"level {l} Scharr derivative differs"


@pytest.mark.parametrize("pair,flags,max_level", [
    ("temporal", 0, 3), ("stereo", 0, 3), ("stereo_prior", 4, 3), ("temporal", 0, 6), ("stereo_prior", 4, 0)])
def test_klt_matches_cv2(gpu_ctx, case, pair, flags, max_level):
    from oracle import klt as oklt
    img0 = case["left"]
    img1 = case["next_left"] if pair == "temporal" else case["right"]
    pts0 = case["pts0"]
    prior = pts0 - np.array([[30.0, 0.0]], np.float32) if flags else None
    gpu_ctx.upload_image(0, img0)
    gpu_ctx.upload_image(1, img1)
    p_g, s_g, e_g = gpu_ctx.klt_track(0, 1, pts0, 21, max_level, flags, prior)
    p_c, s_c, e_c = oklt.lk_cv2(img0, img1, pts0, 21, max_level, flags, prior)
    agree = np.mean(s_g == s_c)
    assert agree >= STATUS_AGREE, f"status agreement {agree}"
    ok = (s_g > 0) & (s_c > 0)
    d = np.abs(p_g - p_c).max(1)[ok]
    frac = np.mean(d <= TOL_PX)
    print(f"{pair} ml={max_level}: ok={ok.sum()} max|dp|={d.max():.3e} p99={np.percentile(d, 99):.3e} within={frac:.5f}")
    assert frac >= STATUS_AGREE
    assert np.percentile(d, 99) < 1e-3
    assert np.abs(e_g - e_c)[ok & (np.abs(p_g - p_c).max(1) < 1e-3)].max() < 0.05
