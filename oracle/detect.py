"""Bucketed corner detector oracle -- TEST INFRASTRUCTURE ONLY.

Restates the bucketing of FeatureExtractor (core/visual_odometry/feature_extractor.h:56-134 WeightBin,
feature_extractor.cpp:94-98 updateWeightBin, :211-282 extractORBwithBinning_fast): bins that already
hold a tracked point get weight 0, every other bin returns its single best-response keypoint, output in
bin-index order.  The keypoint detector itself is cv::ORB in the reference (third-party OpenCV, unpinned,
not under /root/reference); the B200 path defines its own exact-integer Harris response on the resident
Scharr derivative plane (DESIGN.md "K-det"), restated here in numpy int64:
    a = sum_7x7 Ix^2 >> 10, b = sum_7x7 Ix*Iy >> 10, c = sum_7x7 Iy^2 >> 10   (Ix, Iy = cv::Scharr int16)
    score = 25 * (a*c - b*b) - (a+c)^2                                          (Harris, k = 1/25)
candidates: integer pixels with edge <= x < w-edge, edge <= y < h-edge and score > min_score;
ties inside a bin go to the first pixel in raster order (the reference's strict '>' over detection order).
"""
import numpy as np

from . import klt as oklt

f32 = np.float32


def weight_bins(pts, w, h, n_bins_u, n_bins_v):
    """WeightBin::init + reset + update (feature_extractor.h:81-131). Returns (weight[int32], u_step, v_step)."""
    u_step = int(np.floor(f32(w) / f32(n_bins_u)))
    v_step = int(np.floor(f32(h) / f32(n_bins_v)))
    total = n_bins_u * n_bins_v
    weight = np.ones(total, np.int32)
    pts = np.asarray(pts, f32).reshape(-1, 2)
    if len(pts):
        ui = np.floor(pts[:, 0] / f32(u_step)).astype(np.int64)
        vi = np.floor(pts[:, 1] / f32(v_step)).astype(np.int64)
        b = vi * n_bins_u + ui
        b = b[(b >= 0) & (b < total)]
        weight[b] = 0
    return weight, u_step, v_step


def harris_score(img):
    """int64 score plane (h, w) from the bit-exact Scharr derivative of img; 7x7 block."""
    d = oklt.scharr(img).astype(np.int64)          # (h, w, 2)
    ix, iy = d[..., 0], d[..., 1]
    h, w = ix.shape

    def box7(v):
        p = np.zeros((h + 7, w + 7), np.int64)
        p[4:h + 4, 4:w + 4] = v                    # zero outside the image (never used: edge >= 3)
        c = p.cumsum(0).cumsum(1)
        return c[7:, 7:] - c[:-7, 7:] - c[7:, :-7] + c[:-7, :-7]
    a = box7(ix * ix) >> 10
    b = box7(ix * iy) >> 10
    c = box7(iy * iy) >> 10
    return 25 * (a * c - b * b) - (a + c) * (a + c)


def detect_bucketed(img, pts_occupied, n_bins_u, n_bins_v, edge=31, min_score=0):
    """FeatureExtractor::updateWeightBin(pts_occupied) + extractORBwithBinning_fast(img) -> new points [k,2] f32."""
    h, w = img.shape
    weight, u_step, v_step = weight_bins(pts_occupied, w, h, n_bins_u, n_bins_v)
    score = harris_score(img)
    inv_u, inv_v = f32(1.0) / f32(u_step), f32(1.0) / f32(v_step)
    xs = np.arange(w, dtype=f32)
    ys = np.arange(h, dtype=f32)
    ub = np.floor(xs * inv_u).astype(np.int64)     # (int)floor(pt.x * inv_u_step_)
    vb = np.floor(ys * inv_v).astype(np.int64)
    U, V = np.meshgrid(ub, vb)
    X, Y = np.meshgrid(np.arange(w), np.arange(h))
    ok = (X >= edge) & (X < w - edge) & (Y >= edge) & (Y < h - edge) & (U < n_bins_u) & (V < n_bins_v) & (score > min_score)
    B = V * n_bins_u + U
    ok &= weight[np.clip(B, 0, len(weight) - 1)] > 0
    out = []
    if ok.any():
        b, s = B[ok], score[ok]
        r = (Y[ok] * w + X[ok])
        order = np.lexsort((r, -s, b))             # bin asc, score desc, raster asc
        b, r = b[order], r[order]
        first = np.ones(len(b), bool)
        first[1:] = b[1:] != b[:-1]
        for rr in r[first]:
            out.append((rr % w, rr // w))
    return np.asarray(out, f32).reshape(-1, 2)
