"""Mono geometric front-end through the C ABI against the numpy restatements in oracle/mono_step.py:
calcSampsonDistance / calcSymmetricEpipolarDistance (motion_estimator.cpp:539-653) and findInliers1PointHistogram
(:471-537)."""
import numpy as np
import pytest

from oracle import mono_step as omono
from visual_odometry_ros_b200 import capi, synth

pytestmark = pytest.mark.gpu


def test_epipolar_distances_match_restatement(gpu_ctx):
    sc = synth.two_view_scene(seed=21, n=3000)
    R, t = sc["R10"].astype(np.float32), sc["t10"].astype(np.float32)
    F = omono.fundamental(sc["K4"], R, t)
    for symmetric, ref in ((False, omono.sampson(sc["pts0"], sc["pts1"], F)), (True, omono.symmetric_epipolar(sc["pts0"], sc["pts1"], F))):
        g = gpu_ctx.epipolar_distance(sc["pts0"], sc["pts1"], sc["K4"], R, t, symmetric=symmetric)
        # same float32 operation order: equal up to the rounding of the division / sqrt (both IEEE) -> bit-exact expected
        assert np.array_equal(g, ref) or np.max(np.abs(g - ref) / np.maximum(np.abs(ref), 1e-12)) < 1e-6
    gF = gpu_ctx.epipolar_distance(sc["pts0"], sc["pts1"], F10=F)
    assert np.array_equal(gF, gpu_ctx.epipolar_distance(sc["pts0"], sc["pts1"], sc["K4"], R, t))
    inl = np.ones(len(gF), bool)
    inl[sc["outlier_idx"]] = False
    assert np.median(gF[inl]) < 0.5 and np.median(gF[~inl]) > 10.0
    assert len(gpu_ctx.epipolar_distance(sc["pts0"][:0], sc["pts1"][:0], sc["K4"], R, t)) == 0
    with pytest.raises(capi.VoError) as e:
        gpu_ctx.epipolar_distance(sc["pts0"], sc["pts1"][:-1], sc["K4"], R, t)
    assert e.value.status == capi.VO_ERR_SIZE_MISMATCH


@pytest.mark.parametrize("yaw", [0.03, -0.11, 0.0])
def test_one_point_histogram_voting(gpu_ctx, yaw):
    """Planar (yaw + forward) motion: the vote must find the yaw; counts / mask agree with the restatement."""
    sc = synth.two_view_scene(seed=33, n=2500, rotvec=(0.0, yaw, 0.0), t=(np.sin(yaw / 2) * 0.9, 0.0, np.cos(yaw / 2) * 0.9), outlier_frac=0.2)
    g = gpu_ctx.inliers_1point_histogram(sc["pts0"], sc["pts1"], sc["K4"], 5.0)
    th_o, mask_o, theta_o, counts, R_o, t_o = omono.inliers_1point_histogram(sc["pts0"], sc["pts1"], sc["K4"], 5.0)
    # atanf (device) vs numpy: <= 1 ulp apart, so only values on a bin edge can move
    assert np.max(np.abs(g["theta"] - theta_o)) < 1e-6
    top2 = np.sort(counts)[-2:]
    if top2[1] - top2[0] > 4:                      # unambiguous winner
        assert g["theta_opt"] == pytest.approx(th_o, abs=1e-7)
        assert np.abs(g["R10"] - R_o).max() < 1e-6 and np.abs(g["t10"] - t_o).max() < 1e-6
        assert (g["mask"] == mask_o).mean() >= 0.999
    # R10 = Ry(theta) with X1 = R10 X0 + t10: the camera yaw of the scene shows up with the opposite sign, to one bin width
    assert abs(g["theta_opt"] + yaw) <= 1.5 * (1.0 / 400)
    inl = np.ones(len(g["mask"]), bool)
    inl[sc["outlier_idx"]] = False
    assert g["mask"][inl].mean() > 0.9
    e = gpu_ctx.inliers_1point_histogram(sc["pts0"][:0], sc["pts1"][:0], sc["K4"], 5.0)
    assert e["theta_opt"] == pytest.approx(-0.5) and len(e["mask"]) == 0          # empty histogram: the first bin's centre
