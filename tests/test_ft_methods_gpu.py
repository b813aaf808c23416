"""Direct GPU parity of the FeatureTracker entry points -- vo_ft_track, vo_ft_track_with_prior, vo_ft_track_bidirection,
vo_ft_track_bidirection_with_prior, vo_ft_track_with_scale -- against the reference's methods restated over its own
library call (oracle.klt.track* = cv2.calcOpticalFlowPyrLK 4.13 with the argument patterns and post-filters of
core/visual_odometry/feature_tracker.cpp:13-206), on the BASELINE config-2 case (1241x376 stereo pair, 2000 features,
21x21 window, maxLevel 3, thres_err 80, thres_bidirection 0.5), each with a partly-false INCOMING mask:
`mask_valid.resize(n, true)` keeps pre-existing entries (feature_tracker.cpp:20,49,98,178), so
entries that come in false must stay false whatever the tracker finds."""
import numpy as np
import pytest

from visual_odometry_ros_b200 import synth

pytestmark = pytest.mark.gpu

WIN, LVL, THRES_ERR, THRES_BI = 21, 3, 80.0, 0.5      # config/stereo/kitti_00_stereo.yaml:55-59
TOL_PX, AGREE = 0.01, 0.999                           # BASELINE.json north_star


@pytest.fixture(scope="module")
def case():
    c = synth.klt_stereo_case(seed=2002, n=2000)
    rng = np.random.default_rng(77)
    m = np.ones(2000, bool)
    m[rng.choice(2000, 300, replace=False)] = False   # 15 % of the incoming mask is already false
    c["mask_in"] = m
    # a prior like the one StereoVO builds (constant-velocity projection): truth + ~1 px
    from oracle import klt as oklt
    p1, _, _ = oklt.lk_cv2(c["left"], c["next_left"], c["pts0"], WIN, LVL)
    c["prior"] = (p1 + rng.normal(0, 1.0, p1.shape)).astype(np.float32)
    c["scale"] = rng.uniform(0.95, 1.08, 2000).astype(np.float32)
    return c


def _compare(name, p_g, m_g, p_o, m_o, mask_in):
    assert not m_g[~mask_in].any(), f"{name}: entries that came in false must stay false"
    assert not m_o[~mask_in].any()
    agree = float(np.mean(m_g == m_o))
    both = m_g & m_o
    d = np.abs(p_g - p_o).max(1)[both]
    within = float(np.mean(d <= TOL_PX))
    print(f"{name}: valid gpu/oracle {m_g.sum()}/{m_o.sum()} mask agreement {agree:.5f} max|dp| {d.max():.2e} within 0.01 px {within:.5f}")
    assert agree >= AGREE, f"{name}: mask agreement {agree}"
    assert within >= AGREE, f"{name}: positions within 0.01 px {within}"
    assert both.sum() > 1000


def test_ft_track(gpu_ctx, case):
    from oracle import klt as oklt
    gpu_ctx.upload_image(0, case["left"]); gpu_ctx.upload_image(1, case["next_left"])
    p_g, m_g = gpu_ctx.ft_track(0, 1, case["pts0"], WIN, LVL, THRES_ERR, case["mask_in"])
    p_o, m_o = oklt.track(oklt.lk_cv2, case["left"], case["next_left"], case["pts0"], WIN, LVL, THRES_ERR, case["mask_in"])
    _compare("track", p_g, m_g, p_o, m_o, case["mask_in"])


def test_ft_track_with_prior(gpu_ctx, case):
    from oracle import klt as oklt
    gpu_ctx.upload_image(0, case["left"]); gpu_ctx.upload_image(1, case["next_left"])
    p_g, m_g = gpu_ctx.ft_track_with_prior(0, 1, case["pts0"], case["prior"], WIN, LVL, THRES_ERR, case["mask_in"])
    p_o, m_o = oklt.track_with_prior(oklt.lk_cv2, case["left"], case["next_left"], case["pts0"], case["prior"], WIN, LVL, THRES_ERR,
                                     case["mask_in"])
    _compare("trackWithPrior", p_g, m_g, p_o, m_o, case["mask_in"])


def test_ft_track_bidirection(gpu_ctx, case):
    from oracle import klt as oklt
    gpu_ctx.upload_image(0, case["left"]); gpu_ctx.upload_image(1, case["right"])
    p_g, m_g = gpu_ctx.ft_track_bidirection(0, 1, case["pts0"], WIN, LVL, THRES_ERR, THRES_BI, case["mask_in"])
    p_o, m_o = oklt.track_bidirection(oklt.lk_cv2, case["left"], case["right"], case["pts0"], WIN, LVL, THRES_ERR, THRES_BI,
                                      case["mask_in"])
    _compare("trackBidirection", p_g, m_g, p_o, m_o, case["mask_in"])


def test_ft_track_bidirection_with_prior(gpu_ctx, case):
    from oracle import klt as oklt
    gpu_ctx.upload_image(0, case["left"]); gpu_ctx.upload_image(1, case["next_left"])
    p_g, m_g = gpu_ctx.ft_track_bidirection_with_prior(0, 1, case["pts0"], case["prior"], WIN, LVL, THRES_ERR, THRES_BI, case["mask_in"])
    p_o, m_o = oklt.track_bidirection_with_prior(oklt.lk_cv2, case["left"], case["next_left"], case["pts0"], case["prior"], WIN, LVL,
                                                 THRES_ERR, THRES_BI, case["mask_in"])
    _compare("trackBidirectionWithPrior", p_g, m_g, p_o, m_o, case["mask_in"])


def test_ft_track_with_scale(gpu_ctx, case):
    from oracle import klt as oklt
    gpu_ctx.upload_image(0, case["left"]); gpu_ctx.upload_image(1, case["next_left"])
    p_g, m_g = gpu_ctx.ft_track_with_scale(0, 1, case["pts0"], case["scale"], case["prior"], case["mask_in"])
    p_o, m_o = oklt.track_with_scale(case["left"], case["next_left"], case["pts0"], case["scale"], case["prior"], case["mask_in"])
    assert np.array_equal(p_g[~case["mask_in"]], case["prior"][~case["mask_in"]]), "masked-out features must not move"
    _compare("trackWithScale", p_g, m_g, p_o, m_o, case["mask_in"])


@pytest.mark.parametrize("win", [9, 11, 13, 15, 17, 19, 23, 25, 27, 29, 31])
def test_ft_track_every_odd_window(gpu_ctx, case, win):
    """Every odd window 9..31 runs the TMA-staged kernel (no slow-path fallback); same bar as the 21x21 case."""
    from oracle import klt as oklt
    gpu_ctx.upload_image(0, case["left"]); gpu_ctx.upload_image(1, case["next_left"])
    p_g, m_g = gpu_ctx.ft_track(0, 1, case["pts0"], win, LVL, THRES_ERR, case["mask_in"])
    p_o, m_o = oklt.track(oklt.lk_cv2, case["left"], case["next_left"], case["pts0"], win, LVL, THRES_ERR, case["mask_in"])
    _compare(f"track win={win}", p_g, m_g, p_o, m_o, case["mask_in"])
