"""trackWithScale (K7) against its C restatement for assorted image sizes, patch scales and start offsets."""
import sys
import numpy as np
sys.path.insert(0, ".")
from oracle import klt as oklt
from visual_odometry_ros_b200 import capi, synth

rng = np.random.default_rng(9)
ctx = capi.Context(device=0, max_w=1920, max_h=1200, n_slots=2, max_feat=8192)
bad = n = 0
for (w, h) in [(1241, 376), (640, 480), (333, 247), (1920, 1200), (131, 97)]:
    img0 = synth.textured_image(np.random.default_rng(w * 3 + h), w, h)
    img1 = synth.warp_translate_field(img0, 0.8, -0.6)
    ctx.upload_image(0, img0); ctx.upload_image(1, img1)
    for smin, smax in ((1.0, 1.0), (0.8, 1.25), (0.6, 1.6)):
        k = 500
        m = 30       # patches must stay inside for the faithful / intended semantics to coincide (DESIGN, deviations)
        pts0 = np.stack([rng.uniform(m, w - m, k), rng.uniform(m, h - m, k)], 1).astype(np.float32)
        scale = rng.uniform(smin, smax, k).astype(np.float32)
        start = (pts0 + np.array([0.8, -0.6], np.float32) + rng.normal(0, 0.7, (k, 2))).astype(np.float32)
        pg, mg = ctx.ft_track_with_scale(0, 1, pts0, scale, start)
        po, mo = oklt.track_with_scale(img0, img1, pts0, scale, start)
        agree = float(np.mean(mg == mo))
        both = mg & mo
        dmax = float(np.abs(pg - po).max(1)[both].max()) if both.any() else 0.0
        n += 1
        good = agree >= 0.999 and dmax <= 0.01
        bad += not good
        print((w, h), (smin, smax), "mask agreement %.4f" % agree, "accepted %d" % int(mo.sum()), "max |dp| %.2e" % dmax, "" if good else " <-- MISMATCH")
print("cases", n, "bad", bad)
