"""Edge cases of the frame-level entry points: empty inputs, disabled stages, argument errors, and the
StereoVO(mode, yaml) constructor of the reference (stereo_vo.cpp:9-53, 186-285) on a reference-format yaml."""
import ctypes
import os

import numpy as np
import pytest

from visual_odometry_ros_b200 import capi, synth

pytestmark = pytest.mark.gpu

W, H = synth.SMALL_W, synth.SMALL_H

YAML = """%YAML:1.0
flagDoUndistortion: 0 # KITTI-like: already rectified
Camera.left.fx: {fx}
Camera.left.fy: {fy}
Camera.left.cx: {cx}
Camera.left.cy: {cy}
Camera.left.width: {w}
Camera.left.height: {h}
Camera.right.fx: {fx}
Camera.right.fy: {fy}
Camera.right.cx: {cx}
Camera.right.cy: {cy}
Camera.right.width: {w}
Camera.right.height: {h}
T_lr: !!opencv-matrix # this statement is necessary.
  rows: 4
  cols: 4
  dt: f
  data: [1,0,0,0.5371657189, 0,1,0,0,
         0,0,1,0, 0,0,0,1]
feature_tracker.thres_error: 80.0
feature_tracker.thres_bidirection: 0.5
feature_tracker.thres_sampson: 60.0
feature_tracker.window_size: 21
feature_tracker.max_level: 6
feature_extractor.n_features: 2000
feature_extractor.n_bins_u: 32
feature_extractor.n_bins_v: 12
motion_estimator.thres_poseba_error: 3.0 # pixels
keyframe_update.thres_alive_ratio: 0.6
keyframe_update.thres_trans: 2.0 # meters
keyframe_update.thres_rotation: 15.0 # degrees
keyframe_update.n_max_keyframes_in_window: 9
"""


@pytest.fixture(scope="module")
def frames():
    import torch
    dev = "cuda" if torch.cuda.is_available() else "cpu"
    return synth.stereo_sequence(4, W, H, synth.small_K(), seed=3203, device=dev)


def test_detect_argument_errors_and_empty(frames):
    L, R, T = frames
    ctx = capi.Context(device=0, max_w=W, max_h=H, n_slots=2, max_feat=1024)
    with pytest.raises(capi.VoError):
        ctx.detect_bucketed(0, np.zeros((0, 2)), 8, 4)                 # slot has no image yet
    ctx.upload_image(0, L[0])
    with pytest.raises(capi.VoError):
        ctx.detect_bucketed(0, np.zeros((0, 2)), W + 1, 4)             # more bins than pixels
    with pytest.raises(capi.VoError):
        ctx.detect_bucketed(0, np.zeros((0, 2)), 8, 4, edge=1)         # the 7x7 response needs a 3-px border
    flat = np.full((H, W), 77, np.uint8)
    ctx.upload_image(1, flat)
    assert len(ctx.detect_bucketed(1, np.zeros((0, 2)), 16, 8)) == 0   # no gradient -> no candidate beats min_score
    ctx.close()


def test_frame_step_disabled_stages_and_empty_inputs(frames):
    L, R, T = frames
    K, Tlr = synth.small_K(), synth.kitti_T_lr()
    ctx = capi.Context(device=0, max_w=W, max_h=H, n_slots=4, max_feat=4096)
    common = dict(K_l=K, K_r=K, T_lr=Tlr, win=21, max_level=3, thres_err=80.0, thres_poseba=3.0, thres_bi=0.5)
    e2, e3, e0 = np.zeros((0, 2)), np.zeros((0, 3)), np.zeros(0)
    # first frame with the new-feature stage disabled: nothing to do, nothing returned
    g = ctx.stereo_frame_step(-1, 0, 1, L[0], R[0], e2, e2, e3, e0, None, None, n_bins_u=0, n_bins_v=0, new_depth_gate=False, **common)
    assert len(g["index"]) == 0 and len(g["new_l1"]) == 0 and g["n_detected"] == 0
    # first frame with detection: points come back in bin order, inside the edge margin
    g = ctx.stereo_frame_step(-1, 0, 1, L[0], R[0], e2, e2, e3, e0, None, None, n_bins_u=16, n_bins_v=6, new_depth_gate=False, **common)
    assert 0 < len(g["new_l1"]) <= 96
    assert np.all(g["new_l1"][:, 0] >= 31) and np.all(g["new_l1"][:, 1] >= 31)
    # steady state with zero landmarks: the prior pose is returned (the reference would solve on zero points)
    Twp = np.eye(4, dtype=np.float32); dT = np.eye(4, dtype=np.float32); dT[2, 3] = 1.0
    g = ctx.stereo_frame_step(0, 2, 3, L[1], R[1], e2, e2, e3, e0, Twp, dT, n_bins_u=16, n_bins_v=6, **common)
    assert len(g["index"]) == 0 and np.allclose(g["T_wc"], dT) and len(g["new_l1"]) > 0
    # wrong previous slot size
    with pytest.raises(capi.VoError):
        small = capi.Context(device=0, max_w=W, max_h=H, n_slots=4, max_feat=64)
        small.stereo_frame_step(3, 0, 1, L[1], R[1], np.ones((1, 2)), np.ones((1, 2)), np.ones((1, 3)), np.ones(1), Twp, dT,
                                n_bins_u=0, n_bins_v=0, **common)
    ctx.close()


def test_stereo_vo_from_reference_yaml(frames, tmp_path):
    from visual_odometry_ros_b200 import stereo_vo as svo
    L, R, T = frames
    K = synth.small_K()
    y = tmp_path / "stereo.yaml"
    y.write_text(YAML.format(fx=float(K[0]), fy=float(K[1]), cx=float(K[2]), cy=float(K[3]), w=W, h=H))
    vo = svo.StereoVO(yaml_path=str(y))
    # the yaml constructor is the drop-in path: the reference's extractor (cv::ORB restated) at feature_extractor.thres_fastscore
    # and the reference's arithmetic in the pose-only GN (strict-order sums)
    ref = svo.StereoVO(svo.make_parameters(W, H, K, K, synth.kitti_T_lr(), max_level=6, n_bins_u=32, n_bins_v=12, thres_trans=2.0,
                                           detector="orb", thres_fastscore=20, pose_strict=True))
    for k in range(len(L)):
        vo.trackStereoImages(L[k], R[k], 0.1 * k)
        ref.trackStereoImages(L[k], R[k], 0.1 * k)
        assert np.array_equal(vo.pose(), ref.pose())                    # same parameters -> bit-identical run
        assert vo.frame_info()["n_tracked"] == ref.frame_info()["n_tracked"]
    assert vo.frame_info()["ms_total"] > 0
    gt = np.linalg.inv(T[0]) @ T[len(L) - 1]
    assert np.abs(vo.pose()[:3, 3] - gt[:3, 3]).max() < 0.05
    vo.close(); ref.close()
    with pytest.raises(capi.VoError):
        svo.StereoVO(yaml_path=str(tmp_path / "missing.yaml"))
    # flagDoUndistortion: 1 with zero distortion and an already rectified rig: rectification is the identity up to the
    # reference's own conventions (K_rect = mean focal length / image centre, camera.cpp:401-412), so the step runs
    und = tmp_path / "undist.yaml"
    und.write_text(YAML.format(fx=float(K[0]), fy=float(K[1]), cx=float(K[2]), cy=float(K[3]), w=W, h=H).replace("flagDoUndistortion: 0", "flagDoUndistortion: 1"))
    vu = svo.StereoVO(yaml_path=str(und))
    for k in range(len(L)):
        vu.trackStereoImages(L[k], R[k], 0.1 * k)
    assert vu.frame_info()["n_tracked"] > 100
    assert np.abs(vu.pose()[:3, 3] - gt[:3, 3]).max() < 0.1
    vu.close()


def test_new_entry_points_edge_cases():
    """Degenerate inputs of the round's new entry points: loud, specific statuses instead of garbage."""
    ctx = capi.Context(device=0, max_w=W, max_h=H, n_slots=2, max_feat=1024)
    K = synth.small_K()
    # five-point: all correspondences identical -> every minimal sample is degenerate -> no model (VO_ERR_MODE, mono_vo.cpp:590)
    p = np.tile(np.array([[100.0, 80.0]], np.float32), (40, 1))
    with pytest.raises(capi.VoError) as e:
        ctx.pose_5point(p, p, K, 1.0)
    assert e.value.status == capi.VO_ERR_MODE and "calcPose5PointsAlgorithm" in str(e.value)
    # pure rotation (no translation): a model exists, the call must not fail or return NaNs
    sc = synth.two_view_scene(seed=3, n=200, t=(0.0, 0.0, 1e-9), outlier_frac=0.0)
    try:
        r = ctx.pose_5point(sc["pts0"], sc["pts1"], sc["K4"], 1.0)
        assert np.isfinite(r["R10"]).all() and np.isfinite(r["t10"]).all()
    except capi.VoError as err:
        assert err.status == capi.VO_ERR_MODE
    # K-orb on an image whose every level is thinner than twice the 31-px border: no keypoints, no crash
    tiny = np.random.default_rng(0).integers(0, 256, (60, 200), dtype=np.uint8)
    ctx.upload_image(0, tiny)
    P, R, O = ctx.orb_detect(0, 20)
    assert len(P) == 0
    ctx.set_detector("orb", 20)
    assert len(ctx.detect_bucketed(0, np.zeros((0, 2), np.float32), 8, 4)) == 0
    ctx.set_detector("harris")
    # mono init step with fewer than five features
    img = np.random.default_rng(1).integers(0, 256, (H, W), dtype=np.uint8)
    ctx.upload_image(0, img)
    with pytest.raises(capi.VoError) as e:
        ctx.mono_frame_step(0, 1, img, np.array([[50, 50], [60, 60]], np.float32), None, None, None, np.eye(4, dtype=np.float32), None, K,
                            21, 3, 60.0, 0.5, 1000.0, 5.0, False, thres_5p=1.0, init_mode=True)
    assert e.value.status == capi.VO_ERR_INVALID_ARG
    # epipolar distances with a degenerate (zero) fundamental matrix: NaN, like the reference's 0/0, but no error
    d = ctx.epipolar_distance(sc["pts0"], sc["pts1"], F10=np.zeros((3, 3), np.float32))
    assert np.isnan(d).all()
    ctx.close()


def test_mono_vo_rejects_wrong_image_size():
    from visual_odometry_ros_b200 import mono_vo as mvo
    vo = mvo.MonoVO(mvo.make_parameters(W, H, synth.small_K(), max_level=3, n_bins_u=16, n_bins_v=8, D=(0.0,) * 5))
    with pytest.raises(capi.VoError) as e:
        vo.trackImage(np.zeros((H + 2, W), np.uint8))
    assert "same size as the camera model" in str(e.value)
    vo.close()
