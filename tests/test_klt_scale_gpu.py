"""GPU parity of K-klt-scale (FeatureTracker::trackWithScale, feature_tracker.cpp:236-504) against
the FP32 oracle restatement (oracle/klt_scale_oracle.c), through the C ABI."""
import numpy as np
import pytest

from visual_odometry_ros_b200 import synth

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def case():
    from oracle import klt as oklt
    c = synth.klt_stereo_case(seed=2002, n=2000)
    p1, st, err = oklt.lk_cv2(c["left"], c["next_left"], c["pts0"], 21, 3)
    rng = np.random.default_rng(0)
    c["init"] = (p1 + rng.normal(0, 0.4, p1.shape)).astype(np.float32)
    c["scale"] = rng.uniform(0.95, 1.08, len(p1)).astype(np.float32)
    return c


def test_track_with_scale_matches_oracle(gpu_ctx, case):
    from oracle import klt as oklt
    gpu_ctx.upload_image(0, case["left"])
    gpu_ctx.upload_image(1, case["next_left"])
    mask_in = np.ones(2000, bool)
    mask_in[::13] = False     # pre-masked features must be skipped and stay false
    pt_g, m_g = gpu_ctx.ft_track_with_scale(0, 1, case["pts0"], case["scale"], case["init"], mask_in)
    pt_o, m_o = oklt.track_with_scale(case["left"], case["next_left"], case["pts0"], case["scale"], case["init"], mask_in)
    assert not m_g[::13].any()
    assert np.array_equal(pt_g[::13], case["init"][::13]), "masked-out features must not move"
    agree = np.mean(m_g == m_o)
    both = m_g & m_o
    d = np.abs(pt_g - pt_o).max(1)[both]
    print(f"mask agreement {agree:.5f}, valid {both.sum()}, max|dp|={d.max():.3e}, p99={np.percentile(d, 99):.3e}")
    assert agree >= 0.999
    assert np.mean(d <= 0.01) >= 0.999       # BASELINE.json: tracked positions within 0.01 px
    assert np.percentile(d, 99) < 2e-3


def test_track_with_scale_faithful_equals_intended_in_the_interior(case):
    """Documents the reference's stale-buffer defect: it only matters when samples leave the image."""
    from oracle import klt as oklt
    p0 = case["pts0"]
    interior = (p0[:, 0] > 60) & (p0[:, 0] < synth.KITTI_W - 60) & (p0[:, 1] > 40) & (p0[:, 1] < synth.KITTI_H - 40)
    a, ma = oklt.track_with_scale(case["left"], case["next_left"], p0[interior], case["scale"][interior], case["init"][interior])
    b, mb = oklt.track_with_scale(case["left"], case["next_left"], p0[interior], case["scale"][interior], case["init"][interior],
                                  faithful=True)
    assert np.array_equal(a, b) and np.array_equal(ma, mb)


def test_track_with_scale_size_mismatch(gpu_ctx, case):
    from visual_odometry_ros_b200 import capi
    with pytest.raises(capi.VoError):
        gpu_ctx.ft_track_with_scale(0, 1, case["pts0"], case["scale"], case["init"][:10])


def _border_case(seed=5, n=600):
    """Features crowded against the image borders (the 23 x 23 checkerboard of half of them leaves the image), some
    pre-masked, in random order -- the situation in which the reference's never-reset sample buffers matter."""
    from oracle import klt as oklt
    c = synth.klt_stereo_case(seed=2002, n=2000)
    rng = np.random.default_rng(seed)
    W, H = synth.KITTI_W, synth.KITTI_H
    pts = np.empty((n, 2), np.float32)
    kind = rng.integers(0, 6, n)
    for i in range(n):
        k = kind[i]
        if k == 0: pts[i] = (rng.uniform(3, 14), rng.uniform(3, H - 3))                 # left edge
        elif k == 1: pts[i] = (rng.uniform(W - 15, W - 4), rng.uniform(3, H - 3))       # right edge
        elif k == 2: pts[i] = (rng.uniform(3, W - 3), rng.uniform(3, 14))               # top edge
        elif k == 3: pts[i] = (rng.uniform(3, W - 3), rng.uniform(H - 15, H - 4))       # bottom edge
        else: pts[i] = (rng.uniform(40, W - 40), rng.uniform(40, H - 40))               # interior
    p1, st, err = oklt.lk_cv2(c["left"], c["next_left"], pts, 21, 3)
    init = (np.where(st[:, None] > 0, p1, pts) + rng.normal(0, 0.4, pts.shape)).astype(np.float32)
    scale = rng.uniform(0.95, 1.08, n).astype(np.float32)
    mask = rng.uniform(size=n) > 0.1
    return c, pts, init, scale, mask


def test_track_with_scale_faithful_border_mode_matches_the_faithful_oracle(gpu_ctx):
    """vo_set_scale_mode(ctx, 1): samples that leave the image reuse what the previous feature / iteration left in the
    reference's per-sample buffers (feature_tracker.cpp:324-333).  Checked against the sequential restatement with
    faithful=True on a border-heavy case; the default mode is checked against faithful=False on the same case, and the two
    modes must actually differ there."""
    from oracle import klt as oklt
    c, pts, init, scale, mask = _border_case()
    gpu_ctx.upload_image(0, c["left"])
    gpu_ctx.upload_image(1, c["next_left"])
    out = {}
    for faithful in (False, True):
        gpu_ctx.set_scale_mode(faithful)
        try:
            pt_g, m_g = gpu_ctx.ft_track_with_scale(0, 1, pts, scale, init, mask)
        finally:
            gpu_ctx.set_scale_mode(False)
        pt_o, m_o = oklt.track_with_scale(c["left"], c["next_left"], pts, scale, init, mask, faithful=faithful)
        agree = np.mean(m_g == m_o)
        both = m_g & m_o
        d = np.abs(pt_g - pt_o).max(1)[both]
        print(f"faithful={faithful}: mask agreement {agree:.5f}, valid {both.sum()}, max|dp|={d.max():.3e}, within 0.01 px {np.mean(d <= 0.01):.5f}")
        assert agree >= 0.995
        assert np.mean(d <= 0.01) >= 0.995
        assert np.array_equal(pt_g[~mask], init[~mask]) and not m_g[~mask].any()
        out[faithful] = (pt_g, m_g, pt_o, m_o)
    # the defect is visible on this case: the two oracle modes differ, and each GPU mode follows its own oracle mode
    d_modes = np.abs(out[True][2] - out[False][2]).max(1)
    n_diff = int(((d_modes > 0.01) | (out[True][3] != out[False][3])).sum())
    print(f"features on which the reference's stale buffers change the result: {n_diff} of {len(pts)}")
    assert n_diff >= 5
    touched = (d_modes > 0.01) & out[True][1] & out[True][3]
    if touched.any():
        d_right = np.abs(out[True][0] - out[True][2]).max(1)[touched]
        d_wrong = np.abs(out[False][0] - out[True][2]).max(1)[touched]
        assert np.median(d_right) < np.median(d_wrong)
