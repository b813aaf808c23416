"""S2 as a class: the C++ MonoVO (reference API mono_vo.h:235-267, host glue over the C ABI) next to the oracle
composition (oracle/mono_vo.py) on the rendered corridor sequence.

The only non-deterministic stage of the reference is OpenCV's RANSAC.  Sequence parity is therefore run twice:
* the oracle's five-point hook calls the CUDA stage (vo_pose_5point) with the seed the class uses for that frame, so the two
  pipelines see the same model and everything around it (tracking, Sampson gate, landmark / parallax bookkeeping,
  reconstructions, keyframe rule, LBA packing, write-back) is compared frame by frame;
* the oracle uses the reference's own cv2.findEssentialMat: trajectories agree up to the minimal-sample scatter."""
import numpy as np
import pytest

from oracle import mono_vo as omvo
from visual_odometry_ros_b200 import capi, synth

pytestmark = pytest.mark.gpu

W, H = synth.SMALL_W, synth.SMALL_H
NB_U, NB_V = 32, 12
N_FRAMES = 14
SEED = 11


def _rot_angle(Ra, Rb):
    dR = Ra.astype(np.float64) @ Rb.astype(np.float64).T
    return float(np.arcsin(min(1.0, np.linalg.norm(dR - dR.T) / (2.0 * np.sqrt(2.0)))))


@pytest.fixture(scope="module")
def seq():
    import torch
    dev = "cuda" if torch.cuda.is_available() else "cpu"
    return synth.stereo_sequence(N_FRAMES, W, H, synth.small_K(), seed=3103, device=dev)


def _make(seed=SEED, **kw):
    from visual_odometry_ros_b200 import mono_vo as mvo
    return mvo.MonoVO(mvo.make_parameters(W, H, synth.small_K(), max_level=3, n_bins_u=NB_U, n_bins_v=NB_V, thres_translation=2.0,
                                          seed=seed, **kw))


@pytest.mark.parametrize("detector", ["harris", "orb"])
def test_mono_vo_class_next_to_oracle(seq, detector):
    L, _, T = seq
    K = synth.small_K()
    ctx = capi.Context(device=0, max_w=W, max_h=H, n_slots=2, max_feat=4096)

    def fp_gpu(frame_id, p0, p1):
        r = ctx.pose_5point(p0, p1, K, 1.0, seed=SEED + frame_id)
        return True, r["R10"], r["t10"], r["mask"]
    ora = omvo.MonoVOOracle(W, H, K, omvo.default_params(n_bins_u=NB_U, n_bins_v=NB_V, max_level=3, kf_trans=2.0, detector=detector,
                                                         fast_threshold=15), five_point=fp_gpu)
    # strict pose arithmetic (the reference's sequential FP32 sums): the fast mode's FP64 sums make the mono GN stop one iteration
    # apart from the restatement on some frames (4e-4 m on ~150 points), which is a property of that mode, not of this class
    vo = _make(detector=detector, thres_fastscore=15, pose_strict=True)
    same_ids = n_kf = n_lba = 0
    in_step = True            # no borderline feature has flipped between the two LK implementations yet
    for k in range(len(L)):
        Twc_o, info = ora.track(L[k])
        vo.trackImage(L[k], 0.1 * k)
        fi = vo.frame_info()
        ids, pts = vo.tracks()
        Twc_g = vo.pose()
        assert fi["keyframe"] == int(info["keyframe"]), k
        assert fi["used_5point"] == int(info["used_5point"]), k
        jac = len(np.intersect1d(ids, ora.prev.lm_ids)) / max(len(ids), len(ora.prev.lm_ids))
        assert jac >= 0.99, (k, jac)
        in_step = in_step and np.array_equal(ids, ora.prev.lm_ids)
        if np.array_equal(ids, ora.prev.lm_ids):
            same_ids += 1
            d = np.abs(pts - ora.prev.pts).max(1)
            assert np.mean(d <= 0.05) >= 0.98, (k, float(np.mean(d <= 0.05)), float(d.max()))
            assert fi["n_recon"] == info["n_recon"] + info.get("n_recon_kf", 0), (k, fi["n_recon"], info)
            if info["lba"] is not None:
                assert fi["lba_points"] == info["lba"]["n_points"] and fi["lba_obs"] == info["lba"]["n_obs"]
        n_kf += fi["keyframe"]; n_lba += int(fi["lba_points"] > 0)
        dt = np.abs(Twc_g[:3, 3] - Twc_o[:3, 3]).max()
        print(f"frame {k}: kf={fi['keyframe']} n={len(ids)} ids_equal={np.array_equal(ids, ora.prev.lm_ids)} 5pt={fi['used_5point']} "
              f"recon={fi['n_recon']} dt={dt:.2e} lba={fi['lba_points']}/{fi['lba_obs']}")
        # identical landmark sets: the propagated pixel tolerance; after a flip (one landmark more or less among the ~40-200
        # that carry the mono pose) the two runs are two slightly different, equally valid, odometries
        assert dt <= (2e-4 if in_step else 2e-2), k
        assert _rot_angle(Twc_g[:3, :3], Twc_o[:3, :3]) <= (1e-4 if in_step else 2e-3), k
    # the local BA moves keyframe poses afterwards: the refreshed poses agree too (mono_vo.cpp:1184-1185)
    Po = ora.all_poses()
    for j in range(len(L)):
        assert np.abs(vo.frame_pose(j)[:3, 3] - Po[j][:3, 3]).max() <= (2e-4 if in_step else 2e-2), j
    assert same_ids >= 8 and n_kf >= 3 and n_lba >= 2
    assert vo.launch_count > 0
    assert vo.stats_consistent()          # incremental keyframe statistics == the reference's full refresh (mono_vo.cpp:1142-1152)
    vo.close(); ctx.close()


def test_mono_vo_against_reference_ransac_and_truth(seq):
    """Oracle with cv2.findEssentialMat (the reference's call): OpenCV's samples are its own, so the two trajectories are
    compared after removing the scale each run fixed at initialisation (|t| = 1), and both against the rendered truth."""
    L, _, T = seq
    K = synth.small_K()
    ora = omvo.MonoVOOracle(W, H, K, omvo.default_params(n_bins_u=NB_U, n_bins_v=NB_V, max_level=3, kf_trans=2.0))
    vo = _make()
    Pg, Po = [], []
    for k in range(len(L)):
        ora.track(L[k])
        vo.trackImage(L[k], 0.1 * k)
    Pg = np.stack([vo.frame_pose(j) for j in range(len(L))])
    Po = ora.all_poses()
    T0inv = np.linalg.inv(T[0])
    gt = np.stack([T0inv @ T[k] for k in range(len(L))])

    def align_err(P):
        s = np.linalg.norm(gt[1][:3, 3]) / np.linalg.norm(P[1][:3, 3])            # the unit first step fixes the scale
        e = [np.linalg.norm(s * P[k][:3, 3] - gt[k][:3, 3]) / max(1.0, np.linalg.norm(gt[k][:3, 3])) for k in range(len(P))]
        r = [_rot_angle(P[k][:3, :3], gt[k][:3, :3].astype(np.float32)) for k in range(len(P))]
        return max(e), max(r)
    eg, rg = align_err(Pg)
    eo, ro = align_err(Po)
    print(f"relative translation error vs truth: gpu {eg:.3f}, oracle(cv2 RANSAC) {eo:.3f}; rotation gpu {rg:.2e} rad, oracle {ro:.2e} rad")
    assert abs(np.linalg.norm(Pg[1][:3, 3]) - 1.0) < 1e-4                         # mono_vo.cpp:606
    # monocular drift on 14 frames of a 620x188 rendering: a few percent; the CUDA path is not worse than the reference call
    assert eg <= max(1.25 * eo, 0.08) and rg <= max(1.25 * ro, 1.5e-2)
    vo.close()


def test_mono_vo_deterministic_and_yaml(seq, tmp_path):
    L, _, _ = seq
    a, b = _make(), _make()
    for k in range(8):
        a.trackImage(L[k], 0.1 * k)
        b.trackImage(L[k], 0.1 * k)
        assert np.array_equal(a.pose(), b.pose()), k
        ia, pa = a.tracks()
        ib, pb = b.tracks()
        assert np.array_equal(ia, ib) and np.array_equal(pa, pb)
        assert len(np.unique(ia)) == len(ia)
    a.close(); b.close()
    from visual_odometry_ros_b200 import mono_vo as mvo
    K = synth.small_K()
    y = tmp_path / "mono.yaml"
    y.write_text("%YAML:1.0\nflagDoUndistortion: 0\n" + "".join(f"Camera.{n}: {v}\n" for n, v in zip(("fx", "fy", "cx", "cy"), K)) +
                 f"Camera.width: {W}\nCamera.height: {H}\nfeature_tracker.thres_error: 60.0\nfeature_tracker.max_level: 3 # comment\n"
                 "feature_extractor.n_bins_u: 32\nfeature_extractor.n_bins_v: 12\nmotion_estimator.thres_5p_error: 1.0\n"
                 "map_update.thres_parallax: 1.0\nkeyframe_update.thres_translation: 2.0\n")
    c = mvo.MonoVO(yaml_path=str(y))
    # the yaml constructor is the drop-in path: the reference's extractor (default FAST 20) and the reference's arithmetic
    # in the pose-only GN (strict-order sums)
    d = _make(seed=0, detector="orb", thres_fastscore=20, pose_strict=True)
    for k in range(5):
        c.trackImage(L[k], 0.1 * k)
        d.trackImage(L[k], 0.1 * k)
        assert np.array_equal(c.pose(), d.pose()), k
    c.close(); d.close()
    with pytest.raises(capi.VoError):
        mvo.MonoVO(yaml_path=str(tmp_path / "missing.yaml"))


def test_mono_vo_undistortion_path(seq):
    """flagDoUndistortion with zero distortion: the map of camera.cpp:57-87 is the identity to float rounding, which the
    1/32-px quantisation of cv::remap absorbs -- the run must equal the plain one; with real distortion coefficients it runs."""
    L, _, _ = seq
    a, b = _make(), _make(D=(0.0, 0.0, 0.0, 0.0, 0.0))
    for k in range(6):
        a.trackImage(L[k], 0.1 * k)
        b.trackImage(L[k], 0.1 * k)
        assert np.array_equal(a.pose(), b.pose()), k
    a.close(); b.close()
    c = _make(D=(-0.05, 0.01, 1e-4, -1e-4, 0.0))
    for k in range(6):
        c.trackImage(L[k], 0.1 * k)
    assert c.frame_info()["n_tracked"] > 100 and np.isfinite(c.pose()).all()
    c.close()
