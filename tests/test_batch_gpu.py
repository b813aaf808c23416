"""vo_ft_track_batch (host buffers, chunked upload/compute pipeline on three streams) must return exactly what the
one-pair entry points return, whatever the chunking and whether or not the caller's arrays are page-locked."""
import numpy as np
import pytest

from oracle import klt as oklt
from visual_odometry_ros_b200 import capi, synth

pytestmark = pytest.mark.gpu

W, H, N = 640, 192, 300


@pytest.fixture(scope="module")
def pairs():
    rng = np.random.default_rng(11)
    P = 21                                     # chunks of 8 + 13: both compute streams are used
    base = synth.textured_image(rng, W, H)
    lefts, rights, pts = [], [], []
    for i in range(P):
        l = np.ascontiguousarray(np.roll(base, 7 * i, axis=1))
        lefts.append(l)
        rights.append(synth.warp_translate_field(l, 1.5 + 0.1 * i, -0.7))
        pts.append(synth.grid_features(rng, N, W, H, nx=25, ny=12))
    return lefts, rights, np.stack(pts).astype(np.float32)


@pytest.mark.parametrize("pinned", [False, True])
@pytest.mark.parametrize("with_prior", [False, True])
def test_batch_equals_single_calls(pairs, pinned, with_prior):
    import torch
    lefts, rights, pts0 = pairs
    P = len(lefts)
    ctx = capi.Context(device=0, max_w=W, max_h=H, n_slots=2 * P, max_feat=P * N)
    imgs = torch.empty((2 * P, H, W), dtype=torch.uint8)
    if pinned:
        imgs = imgs.pin_memory()
    for i in range(P):
        imgs[i] = torch.from_numpy(lefts[i]); imgs[P + i] = torch.from_numpy(rights[i])
    hn = imgs.numpy()
    s0, s1 = np.arange(P, dtype=np.int32), np.arange(P, 2 * P, dtype=np.int32)
    prior = (pts0 + np.float32([1.0, -0.5])).astype(np.float32)
    mk = (lambda a: torch.from_numpy(a.copy()).pin_memory().numpy()) if pinned else (lambda a: a.copy())
    p0, pt, m = mk(pts0), mk(prior if with_prior else np.zeros_like(pts0)), mk(np.ones((P, N), np.uint8))
    m[3, ::7] = 0                              # pre-existing mask entries are ANDed in (feature_tracker.cpp:20)
    pt_b, m_b = ctx.ft_track_batch(s0, s1, [hn[i].ctypes.data for i in range(P)], [hn[P + i].ctypes.data for i in range(P)], W, H, W,
                                   p0, 21, 3, 80.0, pts_track=pt, mask=m, with_prior=with_prior)
    for i in range(P):
        mi = np.ones(N, np.uint8)
        if i == 3:
            mi[::7] = 0
        if with_prior:
            p_s, m_s = ctx.ft_track_with_prior(int(s0[i]), int(s1[i]), pts0[i], prior[i], 21, 3, 80.0, mask=mi)
        else:
            p_s, m_s = ctx.ft_track(int(s0[i]), int(s1[i]), pts0[i], 21, 3, 80.0, mask=mi)
        assert np.array_equal(np.asarray(m_b[i]).astype(bool), np.asarray(m_s).astype(bool)), i
        assert np.array_equal(pt_b[i], p_s), i
    # and pair 0 against the reference's own library call
    p_c, m_c = (oklt.track_with_prior(oklt.lk_cv2, lefts[0], rights[0], pts0[0], prior[0], 21, 3, 80.0) if with_prior
                else oklt.track(oklt.lk_cv2, lefts[0], rights[0], pts0[0], 21, 3, 80.0))
    ok = m_c & np.asarray(m_b[0]).astype(bool)
    assert np.mean(m_c == np.asarray(m_b[0]).astype(bool)) >= 0.999
    assert np.abs(pt_b[0][ok] - p_c[ok]).max() <= 0.01
    # NULL image pointers keep the slots' images: same answer again
    p0b, ptb, mb = mk(pts0), mk(prior if with_prior else np.zeros_like(pts0)), mk(np.ones((P, N), np.uint8))
    mb[3, ::7] = 0
    pt_b2, m_b2 = ctx.ft_track_batch(s0, s1, None, None, W, H, W, p0b, 21, 3, 80.0, pts_track=ptb, mask=mb, with_prior=with_prior)
    assert np.array_equal(pt_b2, pt_b) and np.array_equal(m_b2, m_b)
    ctx.close()
