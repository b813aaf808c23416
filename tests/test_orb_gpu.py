"""K-orb through the C ABI: the keypoints of cv::ORB::detect (the reference's extractor, feature_extractor.cpp:26-60)
bit-exact against oracle/orb.py -- which tests/test_oracle_orb.py pins bit-exact against cv2.ORB -- and against cv2.ORB
itself, then the reference's bucketing on top (extractORBwithBinning_fast, :211-282)."""
import numpy as np
import pytest

from oracle import orb as oorb
from visual_odometry_ros_b200 import capi, synth

pytestmark = pytest.mark.gpu


def _key(P, R, O):
    return sorted((int(o), np.float32(p[1]), np.float32(p[0]), np.float32(r)) for p, r, o in zip(P, R, O))


@pytest.fixture(scope="module")
def images():
    import torch
    dev = "cuda" if torch.cuda.is_available() else "cpu"
    rng = np.random.default_rng(0)
    Lk, _, _ = synth.stereo_sequence(2, synth.KITTI_W, synth.KITTI_H, synth.kitti_K(), seed=3003, device=dev)
    Ls, _, _ = synth.stereo_sequence(2, synth.SMALL_W, synth.SMALL_H, synth.small_K(), seed=3103, device=dev)
    return {"kitti": Lk[0], "small": Ls[1], "noise": rng.integers(0, 256, (200, 300), dtype=np.uint8),
            "tie_size": rng.integers(0, 256, (97, 129), dtype=np.uint8),     # 129 / 1.2 = 107.5: the level size rounds in double
            "texture": synth.textured_image(np.random.default_rng(3))}


@pytest.mark.parametrize("name,thr", [("kitti", 15), ("kitti", 25), ("small", 20), ("noise", 15), ("texture", 20), ("tie_size", 20)])
def test_orb_keypoints_bit_exact(gpu_ctx, images, name, thr):
    import cv2
    img = images[name]
    gpu_ctx.upload_image(0, img)
    P, R, O = gpu_ctx.orb_detect(0, thr)
    Po, Ro, Oo = oorb.detect(img, thr)
    assert len(P) == len(Po) > 100
    assert _key(P, R, O) == _key(Po, Ro, Oo)
    o = cv2.ORB_create()
    o.setMaxFeatures(10000); o.setScaleFactor(1.2); o.setNLevels(8); o.setEdgeThreshold(31); o.setFirstLevel(0); o.setWTA_K(2)
    o.setScoreType(cv2.ORB_HARRIS_SCORE); o.setPatchSize(31); o.setFastThreshold(thr)
    ref = sorted((k.octave, np.float32(k.pt[1]), np.float32(k.pt[0]), np.float32(k.response)) for k in o.detect(img, None))
    assert ref == _key(P, R, O)                         # the reference's own library call
    print(f"{name} thr {thr}: {len(P)} keypoints, per level {np.bincount(O, minlength=8).tolist()}")


@pytest.mark.parametrize("bins", [(64, 32), (24, 12), (30, 12)])
def test_orb_bucketed_matches_oracle(gpu_ctx, images, bins):
    img = images["kitti"]
    nbu, nbv = bins
    gpu_ctx.upload_image(0, img)
    gpu_ctx.set_detector("orb", 15)
    try:
        first = gpu_ctx.detect_bucketed(0, np.zeros((0, 2), np.float32), nbu, nbv)
        ref = oorb.detect_bucketed(img, np.zeros((0, 2), np.float32), nbu, nbv, 15)
        assert np.array_equal(first, ref) and len(ref) > 0.5 * nbu * nbv
        occ = ref[::3] + np.float32(0.37)               # tracked points sit at sub-pixel positions
        again = gpu_ctx.detect_bucketed(0, occ, nbu, nbv)
        assert np.array_equal(again, oorb.detect_bucketed(img, occ, nbu, nbv, 15))
        assert len(again) < len(first)
    finally:
        gpu_ctx.set_detector("harris")
    # back on K-det: untouched
    from oracle import detect as odet
    assert np.array_equal(gpu_ctx.detect_bucketed(0, np.zeros((0, 2), np.float32), nbu, nbv), odet.detect_bucketed(img, np.zeros((0, 2), np.float32), nbu, nbv))


def test_orb_argument_checks(gpu_ctx, images):
    gpu_ctx.upload_image(0, images["small"])
    with pytest.raises(capi.VoError):
        gpu_ctx.set_detector("orb", 0)
    with pytest.raises(capi.VoError):
        gpu_ctx.orb_detect(0, 20, max_keypoints=10)     # too small for this image
    a = gpu_ctx.orb_detect(0, 20)
    b = gpu_ctx.orb_detect(0, 20)
    assert _key(*a) == _key(*b)


def test_orb_matches_cv2_on_assorted_sizes():
    """Sizes whose level dimensions hit rounding ties, block / noise / smooth content, low and high FAST thresholds."""
    import cv2
    rng = np.random.default_rng(123)
    ctx = capi.Context(device=0, max_w=1920, max_h=1200, n_slots=1, max_feat=1024)
    n = 0
    for (w, h) in [(752, 480), (333, 247), (1920, 1200), (129, 97), (150, 100), (645, 485)]:
        noise = rng.integers(0, 256, (h, w), dtype=np.uint8)
        blocks = (np.kron(rng.integers(0, 2, ((h + 15) // 16, (w + 15) // 16)), np.ones((16, 16)))[:h, :w] * 200 + 20).astype(np.uint8)
        smooth = cv2.GaussianBlur(noise, (0, 0), 3.0)
        for img, thr in ((noise, 60), (np.ascontiguousarray(blocks), 20), (smooth, 5)):
            o = cv2.ORB_create()
            o.setMaxFeatures(10000); o.setScaleFactor(1.2); o.setNLevels(8); o.setEdgeThreshold(31); o.setFirstLevel(0); o.setWTA_K(2)
            o.setScoreType(cv2.ORB_HARRIS_SCORE); o.setPatchSize(31); o.setFastThreshold(thr)
            ref = sorted((k.octave, np.float32(k.pt[1]), np.float32(k.pt[0]), np.float32(k.response)) for k in o.detect(img, None))
            ctx.upload_image(0, img)
            assert _key(*ctx.orb_detect(0, thr, max_keypoints=40000)) == ref, (w, h, thr)
            n += 1
    assert n == 18
    ctx.close()
