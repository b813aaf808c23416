"""CPU checks of the sequence-level oracle (oracle/stereo_vo.py, oracle/detect.py): the restated glue of
StereoVO::trackStereoImages is exercised on a rendered corridor sequence with known ground-truth motion."""
import numpy as np
import pytest

from oracle import detect as odet
from oracle import stereo_vo as osvo
from visual_odometry_ros_b200 import synth

W, H = synth.SMALL_W, synth.SMALL_H


@pytest.fixture(scope="module")
def seq():
    return synth.stereo_sequence(8, W, H, synth.small_K(), seed=3003, device="cpu")


def test_inv4_matches_numpy():
    rng = np.random.default_rng(0)
    for _ in range(20):
        T = np.eye(4)
        T[:3, :3] = synth.so3_exp(rng.normal(0, 0.3, 3))
        T[:3, 3] = rng.normal(0, 2, 3)
        got = osvo.inv4_f32(T.astype(np.float32))
        assert np.abs(got - np.linalg.inv(T)).max() < 5e-6


def test_weight_bins_follow_the_reference_indexing():
    # floor(x / u_step) with u_step = floor(cols / n_bins): x beyond n_bins * u_step aliases into the next row
    # exactly as WeightBin::update does (feature_extractor.h:119-131)
    w, h, nbu, nbv = 1241, 376, 24, 12
    pts = np.array([[10.0, 10.0], [1240.0, 5.0], [600.0, 375.0]], np.float32)
    weight, us, vs = odet.weight_bins(pts, w, h, nbu, nbv)
    assert (us, vs) == (51, 31)
    assert weight[0] == 0
    assert weight[0 * nbu + 24] == 0            # x = 1240 -> u index 24 == n_bins_u -> aliases to bin (row 1, col 0)
    assert weight.sum() == nbu * nbv - 2        # y = 375 -> v index 12 -> bin index out of range, ignored


def test_detector_picks_one_maximum_per_free_bin(seq):
    L, R, T = seq
    img = L[0]
    nbu, nbv = 32, 12
    score = odet.harris_score(img)
    pts = odet.detect_bucketed(img, np.zeros((0, 2)), nbu, nbv, 31, 0)
    assert len(pts) > 100
    us, vs = W // nbu, H // nbv
    bins = (np.floor(pts[:, 1] / vs) * nbu + np.floor(pts[:, 0] / us)).astype(int)
    assert len(np.unique(bins)) == len(bins) and np.all(np.diff(bins) > 0)        # one per bin, bin-index order
    for (x, y), b in zip(pts.astype(int), bins):
        assert 31 <= x < W - 31 and 31 <= y < H - 31
        v, u = divmod(b, nbu)
        ys, xs = np.mgrid[max(31, v * vs):min(H - 31, (v + 1) * vs), max(31, u * us):min(W - 31, (u + 1) * us)]
        assert score[y, x] == score[ys, xs].max()
    # occupied bins return nothing
    pts2 = odet.detect_bucketed(img, pts, nbu, nbv, 31, 0)
    bins2 = (np.floor(pts2[:, 1] / vs) * nbu + np.floor(pts2[:, 0] / us)).astype(int)
    assert len(np.intersect1d(bins, bins2)) == 0


def test_sequence_oracle_tracks_the_ground_truth(seq):
    L, R, T = seq
    K, Tlr = synth.small_K(), synth.kitti_T_lr()
    vo = osvo.StereoVOOracle(W, H, K, K, Tlr, osvo.default_params(n_bins_u=32, n_bins_v=12, kf_trans=2.0))
    T0inv = np.linalg.inv(T[0])
    n_kf = n_lba = 0
    for k in range(len(L)):
        Twc, info = vo.track(L[k], R[k])
        gt = T0inv @ T[k]
        assert np.abs(Twc[:3, 3] - gt[:3, 3]).max() <= 0.02 + 0.01 * np.linalg.norm(gt[:3, 3]), k
        n_kf += int(info["keyframe"])
        if info["lba"] is not None:
            n_lba += 1
            assert info["lba"]["n_points"] > 100 and info["lba"]["ok"]
            assert info["lba"]["avg_err"][-1] <= info["lba"]["avg_err"][0] + 1e-9
        if k == 0:
            assert not info["keyframe"] and info["n_recon"] > 100        # the first frame never becomes a keyframe
        if k == 1:
            assert info["keyframe"]                                      # the second always does (empty window)
    assert n_kf >= 3 and n_lba >= 1
    # landmark bookkeeping invariants
    ids = vo.prev.lm_ids
    assert len(np.unique(ids)) == len(ids)
    assert all(vo.last_frame[i] == vo.prev.id for i in ids)
    for i in ids[:50]:
        obs = vo.kf_obs[i]
        assert all(obs[j][1] == j % 2 for j in range(len(obs)))          # L then R per keyframe (keyframes.cpp:196-214)
