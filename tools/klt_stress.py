"""Pyramids / Scharr (bit-exact) and LK (0.01 px, status) against cv2 on assorted image sizes, windows and level counts."""
import sys
import numpy as np
sys.path.insert(0, ".")
from oracle import klt as oklt
from visual_odometry_ros_b200 import capi, synth

rng = np.random.default_rng(77)
ctx = capi.Context(device=0, max_w=1920, max_h=1200, n_slots=2, max_feat=8192)
bad = 0
n = 0
for (w, h) in [(1241, 376), (640, 480), (333, 247), (131, 77), (1920, 1200), (752, 480), (97, 129), (65, 64), (1000, 61)]:
    img0 = synth.textured_image(np.random.default_rng(w + h), w, h)
    img1 = synth.warp_translate_field(img0, 1.7, -0.9)
    for win, lvl in ((21, 3), (15, 2), (13, 4), (9, 1), (31, 3)):
        m = win
        if w <= 2 * m + 4 or h <= 2 * m + 4:
            continue
        k = 400
        pts = np.stack([rng.uniform(m, w - m, k), rng.uniform(m, h - m, k)], 1).astype(np.float32)
        ctx.upload_image(0, img0); ctx.upload_image(1, img1)
        pg, sg, eg = ctx.klt_track(0, 1, pts, win, lvl)
        pc, sc, ec = oklt.lk_cv2(img0, img1, pts, win, lvl)
        ok = (sg > 0) & (sc > 0)
        st = float(np.mean(sg == sc))
        dmax = float(np.abs(pg - pc).max(1)[ok].max()) if ok.any() else 0.0
        lv, dv = oklt.build_pyramid(img0, win, lvl)
        pyr_ok = True
        for l in range(len(lv)):
            gi, gd = ctx.read_pyramid_level(0, l)
            pyr_ok &= np.array_equal(gi, lv[l]) and np.array_equal(gd, dv[l])
        n += 1
        good = pyr_ok and st >= 0.999 and dmax <= 0.01
        bad += not good
        if not good:
            print("MISMATCH", (w, h), win, lvl, "pyr", pyr_ok, "status", st, "dmax", dmax)
print("cases", n, "bad", bad)
