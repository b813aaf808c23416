"""CPU: pin the KLT oracle (oracle/klt_oracle.c, oracle/klt_scale_oracle.c) against the reference's own
library (cv2 4.13 == cv::calcOpticalFlowPyrLK / buildOpticalFlowPyramid / Sobel, the calls made at
feature_tracker.cpp:29..186 and stereo_vo.cpp:551-552) -- live and against the committed golden vectors."""
import os

import numpy as np
import pytest

from oracle import klt as oklt
from visual_odometry_ros_b200 import synth

GOLD = np.load(os.path.join(os.path.dirname(__file__), "golden", "klt_cv2_golden.npz"))


@pytest.fixture(scope="module")
def small_case():
    return synth.klt_stereo_case(seed=31, n=300, w=420, h=200)


def test_pyrdown_scharr_bit_exact_vs_cv2_live(small_case):
    import cv2
    x = small_case["left"]
    for _ in range(3):
        assert np.array_equal(oklt.pyrdown(x), cv2.pyrDown(x))
        s = oklt.scharr(x)
        assert np.array_equal(s[..., 0], cv2.Scharr(x, cv2.CV_16S, 1, 0))
        assert np.array_equal(s[..., 1], cv2.Scharr(x, cv2.CV_16S, 0, 1))
        x = cv2.pyrDown(x)


@pytest.mark.parametrize("w,h", [(64, 48), (67, 45), (33, 97)])
def test_pyrdown_odd_sizes(w, h):
    import cv2
    img = np.random.default_rng(w * h).integers(0, 256, (h, w)).astype(np.uint8)
    assert np.array_equal(oklt.pyrdown(img), cv2.pyrDown(img))
    du, dv = oklt.sobel3_f32(img)
    assert np.array_equal(du, cv2.Sobel(img, cv2.CV_32F, 1, 0, ksize=3))
    assert np.array_equal(dv, cv2.Sobel(img, cv2.CV_32F, 0, 1, ksize=3))


def test_pyramid_matches_golden_and_level_clamp():
    lv, dv = oklt.build_pyramid(GOLD["img0"], 21, 3)
    assert len(lv) == int(GOLD["pyr_nlevels"])
    for l in range(len(lv)):
        assert np.array_equal(lv[l], GOLD[f"pyr_img{l}"])
        assert np.array_equal(dv[l], GOLD[f"pyr_der{l}"])
    # OpenCV's clamp: 1241x376 / win 21 / maxLevel 6 -> 4
    assert oklt.effective_max_level(1241, 376, 21, 6) == 4
    assert oklt.effective_max_level(1241, 376, 21, 3) == 3
    assert oklt.effective_max_level(40, 40, 21, 3) == 0


def test_sobel_matches_golden():
    du, dv = oklt.sobel3_f32(GOLD["img0"])
    assert np.array_equal(du, GOLD["sobel_du"]) and np.array_equal(dv, GOLD["sobel_dv"])


@pytest.mark.parametrize("win", [21, 15])
@pytest.mark.parametrize("ml", [0, 2, 6])
@pytest.mark.parametrize("mode", ["fwd", "pri"])
def test_lk_restatement_matches_golden_cv2(win, ml, mode):
    flags = 4 if mode == "pri" else 0
    p, s, e = oklt.lk_c(GOLD["img0"], GOLD["img1"], GOLD["pts0"], win, ml, flags, GOLD["prior"])
    gp, gs, ge = GOLD[f"{mode}_w{win}_l{ml}_p"], GOLD[f"{mode}_w{win}_l{ml}_s"], GOLD[f"{mode}_w{win}_l{ml}_e"]
    assert np.mean(s == gs) >= 0.99
    ok = (s > 0) & (gs > 0)
    d = np.abs(p - gp).max(1)[ok]
    assert np.percentile(d, 99) < 2e-3 and np.mean(d <= 0.01) >= 0.99
    assert np.abs(e - ge)[ok][d < 1e-3].max() < 0.05


def test_lk_restatement_matches_cv2_live(small_case):
    c = small_case
    for flags, prior in ((0, None), (4, c["pts0"] - np.array([[20, 0]], np.float32))):
        p, s, e = oklt.lk_c(c["left"], c["right"], c["pts0"], 21, 3, flags, prior)
        gp, gs, ge = oklt.lk_cv2(c["left"], c["right"], c["pts0"], 21, 3, flags, prior)
        assert np.mean(s == gs) >= 0.999
        ok = (s > 0) & (gs > 0)
        assert np.percentile(np.abs(p - gp).max(1)[ok], 99) < 1e-3


@pytest.mark.parametrize("name", ["track", "track_with_prior", "track_bidirection", "track_bidirection_with_prior"])
def test_feature_tracker_front_ends_match_golden(name):
    """The four FeatureTracker post-filters (feature_tracker.cpp:13-206) on top of the C restatement."""
    args = dict(track=(21, 3, 30.0), track_with_prior=(GOLD["prior"], 21, 3, 30.0),
                track_bidirection=(21, 3, 30.0, 0.5), track_bidirection_with_prior=(GOLD["prior"], 21, 3, 30.0, 0.5))[name]
    p, m = getattr(oklt, name)(oklt.lk_c, GOLD["img0"], GOLD["img1"], GOLD["pts0"], *args)
    assert np.mean(m == GOLD[f"ft_{name}_m"]) >= 0.98
    both = m & GOLD[f"ft_{name}_m"]
    assert np.abs(p - GOLD[f"ft_{name}_p"]).max(1)[both].max() < 0.01


def test_mask_in_is_anded_not_overwritten():
    """mask_valid.resize(n, true) keeps pre-existing entries (Appendix B #7)."""
    m_in = np.ones(len(GOLD["pts0"]), bool)
    m_in[::3] = False
    _, m = oklt.track(oklt.lk_c, GOLD["img0"], GOLD["img1"], GOLD["pts0"], 21, 3, 30.0, mask_in=m_in)
    assert not m[::3].any()


def test_empty_inputs():
    p, s, e = oklt.lk_c(GOLD["img0"], GOLD["img1"], np.zeros((0, 2), np.float32), 21, 3)
    assert p.shape == (0, 2) and len(s) == 0


def test_track_with_scale_regression():
    reg = np.load(os.path.join(os.path.dirname(__file__), "golden", "oracle_regression.npz"))
    pt, m = oklt.track_with_scale(GOLD["img0"], GOLD["img1"], GOLD["pts0"], np.full(len(GOLD["pts0"]), 1.03, np.float32),
                                  GOLD["prior"])
    assert np.array_equal(m, reg["kscale_m"]) and np.allclose(pt, reg["kscale_p"], atol=1e-4)
    # pure translation, scale 1: the refinement must land on the true shift (1.7, -0.9)
    pt1, m1 = oklt.track_with_scale(GOLD["img0"], GOLD["img1"], GOLD["pts0"], np.ones(len(GOLD["pts0"]), np.float32),
                                    GOLD["prior"])
    inner = m1 & (GOLD["pts0"][:, 0] > 30) & (GOLD["pts0"][:, 0] < 170) & (GOLD["pts0"][:, 1] > 30) & (GOLD["pts0"][:, 1] < 106)
    flow = (pt1 - GOLD["pts0"])[inner]
    assert np.abs(np.median(flow, 0) - np.array([1.7, -0.9])).max() < 0.1
