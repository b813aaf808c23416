"""cv::ORB keypoint detection oracle -- TEST INFRASTRUCTURE ONLY.

The reference detects with ``cv::ORB`` configured at core/visual_odometry/feature_extractor.cpp:26-60 (maxFeatures 10000,
scaleFactor 1.2, 8 levels, edgeThreshold 31, HARRIS_SCORE, patchSize 31, fastThreshold from the yaml) and buckets the
keypoints in ``extractORBwithBinning_fast`` (:211-282).  OpenCV is third-party and not under /root/reference, so this file
restates the published algorithm of ``cv::ORB::detect`` (OpenCV 4.x ``features2d/src/orb.cpp``: ``computeKeyPoints``,
``HarrisResponses``; ``fast.cpp``: ``FAST_t<16>`` + ``cornerScore<16>``; ``keypoint.cpp``: ``runByImageBorder``,
``retainBest``; ``imgproc/src/resize.cpp``: ``INTER_LINEAR_EXACT`` with 8.8 fixed-point coefficients) in numpy, stage by
stage, and ``tests/test_oracle_orb.py`` pins every stage against the cv2 4.13 wheel -- the very library the reference links:
``cv2.resize(INTER_LINEAR_EXACT)``, ``cv2.FastFeatureDetector``, ``cv2.ORB.detect``.

``detect`` returns the keypoints of cv::ORB::detect as (pt [n,2] float32 in level-0 pixels, response [n] float32,
octave [n]) in a canonical order (level, then raster); OpenCV's own order inside a level is whatever ``std::nth_element``
leaves, which only matters to the reference when two keypoints of one bucket have bit-equal Harris responses.
"""
import numpy as np

f32 = np.float32
HARRIS_K = f32(0.04)

# 16-pixel Bresenham circle of radius 3 (fast.cpp makeOffsets), starting at (0, 3) going clockwise in image coordinates
_RING = [(0, 3), (1, 3), (2, 2), (3, 1), (3, 0), (3, -1), (2, -2), (1, -3), (0, -3), (-1, -3), (-2, -2), (-3, -1), (-3, 0), (-3, 1),
         (-2, 2), (-1, 3)]


def level_sizes(w, h, n_levels=8, scale_factor=1.2):
    """orb.cpp detectAndCompute: the level size is cvRound(dim / scale) with the DOUBLE scale pow(scaleFactor, level)
    (pinned on 129 x 97: 129 / 1.2 = 107.5 -> 108, where the float scale 1.2f would give 107); the keypoint coordinates are
    multiplied by layerScale = (float)pow(scaleFactor, level)."""
    out = []
    for lv in range(n_levels):
        sd = np.power(np.float64(scale_factor), np.float64(lv))
        out.append((int(np.rint(np.float64(w) / sd)), int(np.rint(np.float64(h) / sd)), f32(sd)))
    return out


def _coeffs(src, dst):
    """interpolationLinear<uchar>::getCoeffs: offsets and 8.8 fixed-point weights of one axis."""
    scale = np.float64(1.0) / (np.float64(dst) / np.float64(src))
    d = np.arange(dst, dtype=np.float64)
    fval = scale * (d + 0.5) - 0.5
    ival = np.floor(fval).astype(np.int64)
    ofs = np.clip(ival, 0, src - 1)
    c1 = np.rint((fval - ival) * 256.0).astype(np.int64)          # cvRound: half to even
    left = (ival < 0) | (src <= 1)
    right = (~left) & (ival >= src - 1)
    c1 = np.where(left | right, 0, c1)
    ofs1 = np.where(left | right, ofs, np.minimum(ofs + 1, src - 1))
    return ofs, ofs1, c1


def resize_linear_exact(img, dw, dh):
    """cv::resize(INTER_LINEAR_EXACT) for CV_8UC1 (resize_bitExact<uchar, interpolationLinear>)."""
    h, w = img.shape
    xo, xo1, xc = _coeffs(w, dw)
    yo, yo1, yc = _coeffs(h, dh)
    s = img.astype(np.int64)
    hl = np.minimum(s[:, xo] * (256 - xc) + s[:, xo1] * xc, 65535)              # ufixedpoint16
    v = np.minimum(hl[yo, :] * (256 - yc)[:, None] + hl[yo1, :] * yc[:, None], 0xFFFFFFFF)
    return np.minimum((v + 32768) >> 16, 255).astype(np.uint8)


def pyramid(img, n_levels=8, scale_factor=1.2):
    h, w = img.shape
    out = [img]
    for lv, (lw, lh, _) in enumerate(level_sizes(w, h, n_levels, scale_factor)):
        if lv:
            out.append(resize_linear_exact(out[-1], lw, lh))
    return out


def fast_scores(img, threshold):
    """FAST-9/16 corner test + cornerScore<16> for every pixel (score 0 where not a corner or within 3 px of the border)."""
    h, w = img.shape
    v = img.astype(np.int32)
    c = v[3:h - 3, 3:w - 3]
    d = np.stack([c - v[3 + dy:h - 3 + dy, 3 + dx:w - 3 + dx] for dx, dy in _RING], 0)        # d[k] = v - ring[k]
    d = np.concatenate([d, d[:9]], 0)                                                     # 25 entries, wrap
    # best arc of 9 contiguous ring pixels: all darker than the centre by more than t (d > t) or all brighter (d < -t)
    amin = np.full(c.shape, -1000, np.int32)
    bmax = np.full(c.shape, 1000, np.int32)
    for k in range(16):
        arc = d[k:k + 9]
        amin = np.maximum(amin, arc.min(0))
        bmax = np.minimum(bmax, arc.max(0))
    corner = (amin > threshold) | (bmax < -threshold)
    score = np.maximum(np.maximum(amin, -bmax), threshold) - 1
    out = np.zeros((h, w), np.int32)
    out[3:h - 3, 3:w - 3] = np.where(corner, score, 0)
    return out


def fast_detect(img, threshold):
    """cv::FastFeatureDetector(threshold, nonmaxSuppression = true, TYPE_9_16): (x, y, score) in raster order."""
    s = fast_scores(img, threshold)
    h, w = s.shape
    p = np.pad(s, 1)
    nb = np.stack([p[1 + dy:1 + dy + h, 1 + dx:1 + dx + w] for dy in (-1, 0, 1) for dx in (-1, 0, 1) if (dx, dy) != (0, 0)], 0)
    keep = (s > 0) & (s > nb.max(0))
    # fast.cpp: rows 3 .. rows-4,
    # columns 3 .. cols-4
    keep[:3] = False; keep[h - 3:] = False; keep[:, :3] = False; keep[:, w - 3:] = False
    ys, xs = np.nonzero(keep)
    return xs.astype(np.int32), ys.astype(np.int32), s[ys, xs]


def retain_best_mask(resp, n):
    """KeyPointsFilter::retainBest as a set: everything with response >= the n-th largest (ties kept)."""
    if n < 0 or len(resp) <= n:
        return np.ones(len(resp), bool)
    if n == 0:
        return np.zeros(len(resp), bool)
    cut = np.partition(resp, len(resp) - n)[len(resp) - n]
    return resp >= cut


def harris_responses(img, xs, ys, block=7):
    """orb.cpp HarrisResponses: 7x7 block of Sobel-like integer gradients, float32 response."""
    v = img.astype(np.int64)
    r = block // 2
    a = np.zeros(len(xs), np.int64); b = np.zeros(len(xs), np.int64); c = np.zeros(len(xs), np.int64)
    for dy in range(-r, block - r):
        for dx in range(-r, block - r):
            y, x = ys + dy, xs + dx
            ix = (v[y, x + 1] - v[y, x - 1]) * 2 + (v[y - 1, x + 1] - v[y - 1, x - 1]) + (v[y + 1, x + 1] - v[y + 1, x - 1])
            iy = (v[y + 1, x] - v[y - 1, x]) * 2 + (v[y + 1, x - 1] - v[y - 1, x - 1]) + (v[y + 1, x + 1] - v[y - 1, x + 1])
            a += ix * ix; b += iy * iy; c += ix * iy
    scale = f32(f32(1.0) / f32(f32(4 * block) * f32(255.0)))
    s4 = f32(f32(f32(scale * scale) * scale) * scale)
    af, bf, cf = a.astype(f32), b.astype(f32), c.astype(f32)          # int -> float (a, b, c fit int32 in OpenCV)
    t0 = (af * bf).astype(f32)
    t1 = (cf * cf).astype(f32)
    sab = (af + bf).astype(f32)
    t2 = ((HARRIS_K * sab).astype(f32) * sab).astype(f32)
    return (((t0 - t1).astype(f32) - t2).astype(f32) * s4).astype(f32)


def features_per_level(nfeatures=10000, n_levels=8, scale_factor=1.2):
    factor = f32(1.0 / scale_factor)
    nd = f32(f32(nfeatures) * f32(f32(1) - factor) / f32(f32(1) - f32(np.power(np.float64(factor), np.float64(n_levels)))))
    out, total = [], 0
    for _ in range(n_levels - 1):
        out.append(int(np.rint(nd)))
        total += out[-1]
        nd = f32(nd * factor)
    out.append(max(nfeatures - total, 0))
    return out


def detect(img, fast_threshold, nfeatures=10000, n_levels=8, scale_factor=1.2, edge=31):
    """cv::ORB::detect (HARRIS_SCORE) -> pt [n,2] f32 (level-0 pixels), response [n] f32, octave [n] int32."""
    h, w = img.shape
    pyr = pyramid(img, n_levels, scale_factor)
    sizes = level_sizes(w, h, n_levels, scale_factor)
    quota = features_per_level(nfeatures, n_levels, scale_factor)
    P, R, O = [], [], []
    for lv, im in enumerate(pyr):
        lh, lw = im.shape
        if lw <= 2 * edge or lh <= 2 * edge:
            continue
        xs, ys, sc = fast_detect(im, fast_threshold)
        inb = (xs >= edge) & (xs < lw - edge) & (ys >= edge) & (ys < lh - edge)          # runByImageBorder
        xs, ys, sc = xs[inb], ys[inb], sc[inb]
        k = retain_best_mask(sc.astype(f32), 2 * quota[lv])
        xs, ys = xs[k], ys[k]
        hr = harris_responses(im, xs, ys)
        k = retain_best_mask(hr, quota[lv])
        xs, ys, hr = xs[k], ys[k], hr[k]
        s = sizes[lv][2]
        P.append(np.stack([xs.astype(f32) * s, ys.astype(f32) * s], 1).astype(f32))
        R.append(hr); O.append(np.full(len(xs), lv, np.int32))
    if not P:
        return np.zeros((0, 2), f32), np.zeros(0, f32), np.zeros(0, np.int32)
    return np.concatenate(P), np.concatenate(R), np.concatenate(O)


_CV_ORB = {}


def detect_cv2(img, fast_threshold):
    """The reference's own library call: cv::ORB configured as feature_extractor.cpp:30-56, keypoints in OpenCV's order."""
    import cv2
    o = _CV_ORB.get(int(fast_threshold))
    if o is None:
        o = cv2.ORB_create()
        o.setMaxFeatures(10000); o.setScaleFactor(1.2); o.setNLevels(8); o.setEdgeThreshold(31); o.setFirstLevel(0); o.setWTA_K(2)
        o.setScoreType(cv2.ORB_HARRIS_SCORE); o.setPatchSize(31); o.setFastThreshold(int(fast_threshold))
        _CV_ORB[int(fast_threshold)] = o
    kp = o.detect(np.ascontiguousarray(img), None)
    if not kp:
        return np.zeros((0, 2), f32), np.zeros(0, f32), np.zeros(0, np.int32)
    return (np.asarray([k.pt for k in kp], f32), np.asarray([k.response for k in kp], f32), np.asarray([k.octave for k in kp], np.int32))


def detect_bucketed(img, occupied, n_bins_u, n_bins_v, fast_threshold, backend="numpy", **kw):
    """FeatureExtractor::updateWeightBin + extractORBwithBinning_fast (feature_extractor.cpp:94-98, 211-282) over
    cv::ORB keypoints: the best-response keypoint of every bin that holds no occupied point, in bin order.
    backend "numpy": the restatement above; "cv2": cv2.ORB itself (same keypoint set, pinned; much faster)."""
    from .detect import weight_bins
    h, w = img.shape
    pts, resp, _ = detect_cv2(img, fast_threshold) if backend == "cv2" else detect(img, fast_threshold, **kw)
    weight, u_step, v_step = weight_bins(occupied, w, h, n_bins_u, n_bins_v)
    inv_u, inv_v = f32(1.0) / f32(u_step), f32(1.0) / f32(v_step)
    best = np.full(n_bins_u * n_bins_v, -1, np.int64)
    score = np.full(n_bins_u * n_bins_v, f32(-1), f32)
    if len(pts):
        u = np.floor((pts[:, 0] * inv_u).astype(f32)).astype(np.int64)
        v = np.floor((pts[:, 1] * inv_v).astype(f32)).astype(np.int64)
        for i in range(len(pts)):
            if not (0 <= u[i] < n_bins_u and 0 <= v[i] < n_bins_v):
                continue
            b = v[i] * n_bins_u + u[i]
            if weight[b] and score[b] < resp[i]:
                score[b] = resp[i]; best[b] = i
    sel = best[best >= 0]
    return pts[sel].astype(f32)
