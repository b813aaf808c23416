"""Pins oracle/orb.py -- the numpy restatement of cv::ORB::detect as the reference configures it
(core/visual_odometry/feature_extractor.cpp:26-60) -- against the cv2 4.13 wheel, stage by stage: the INTER_LINEAR_EXACT
pyramid, the FAST-9/16 keypoints and scores of every tested level, and the final keypoint set (octave, point, Harris
response) of cv2.ORB.detect, all bit-exact."""
import numpy as np
import pytest

from oracle import orb
from visual_odometry_ros_b200 import synth

cv2 = pytest.importorskip("cv2")


def reference_orb(thr):
    o = cv2.ORB_create()            # feature_extractor.cpp:30, then the setters of :48-56 (setScaleFactor takes a double)
    o.setMaxFeatures(10000); o.setScaleFactor(1.2); o.setNLevels(8); o.setEdgeThreshold(31); o.setFirstLevel(0); o.setWTA_K(2)
    o.setScoreType(cv2.ORB_HARRIS_SCORE); o.setPatchSize(31); o.setFastThreshold(thr)
    return o


def images():
    rng = np.random.default_rng(0)
    L, _, _ = synth.stereo_sequence(2, synth.SMALL_W, synth.SMALL_H, synth.small_K(), seed=3103, device="cpu")
    return {"corridor": L[0], "noise": rng.integers(0, 256, (200, 300), dtype=np.uint8),
            "tie_size": rng.integers(0, 256, (97, 129), dtype=np.uint8),     # 129 / 1.2 = 107.5: the level size rounds in double
            "texture": np.ascontiguousarray(synth.textured_image(np.random.default_rng(3))[:300, :700])}


@pytest.mark.parametrize("name", ["corridor", "noise", "texture", "tie_size"])
def test_orb_stages_bit_exact_against_cv2(name):
    img = images()[name]
    h, w = img.shape
    pyr = orb.pyramid(img)
    prev = img
    for lv, (lw, lh, _) in enumerate(orb.level_sizes(w, h)):
        if lv:
            prev = cv2.resize(prev, (lw, lh), interpolation=cv2.INTER_LINEAR_EXACT)
            assert np.array_equal(prev, pyr[lv]), lv
    for thr in (15, 20):
        fd = cv2.FastFeatureDetector_create(thr, True, cv2.FAST_FEATURE_DETECTOR_TYPE_9_16)
        for lv in (0, 2, 4):
            ref = sorted((int(k.pt[1]), int(k.pt[0]), int(k.response)) for k in fd.detect(pyr[lv]))
            xs, ys, sc = orb.fast_detect(pyr[lv], thr)
            assert ref == sorted(zip(ys.tolist(), xs.tolist(), sc.tolist())), (thr, lv)
        ref = sorted((k.octave, np.float32(k.pt[1]), np.float32(k.pt[0]), np.float32(k.response)) for k in reference_orb(thr).detect(img, None))
        P, R, O = orb.detect(img, thr)
        got = sorted((int(o), np.float32(p[1]), np.float32(p[0]), np.float32(r)) for p, r, o in zip(P, R, O))
        assert len(ref) > 100 and ref == got, (thr, len(ref), len(got))


def test_orb_bucketing_follows_reference():
    img = images()["corridor"]
    h, w = img.shape
    pts_all = orb.detect_bucketed(img, np.zeros((0, 2), np.float32), 32, 12, 15)
    assert 100 < len(pts_all) <= 32 * 12
    occ = pts_all[::2]
    pts = orb.detect_bucketed(img, occ, 32, 12, 15)
    ub, vb = w // 32, h // 12
    occ_bins = set((int(p[1] // vb) * 32 + int(p[0] // ub)) for p in occ)
    got_bins = [int(p[1] // vb) * 32 + int(p[0] // ub) for p in pts]
    assert not (set(got_bins) & occ_bins) and got_bins == sorted(got_bins) and len(set(got_bins)) == len(got_bins)
    # the cv2 backend (the reference's own call, OpenCV's keypoint order) buckets to the same points
    assert np.array_equal(pts, orb.detect_bucketed(img, occ, 32, 12, 15, backend="cv2"))
    assert np.array_equal(pts_all, orb.detect_bucketed(img, np.zeros((0, 2), np.float32), 32, 12, 15, backend="cv2"))
