"""Bucketed extraction in ORB mode against the reference composition (cv2.ORB keypoints in OpenCV's order + the
restated WeightBin bucketing) for assorted sizes, bin layouts and occupancies."""
import sys
import numpy as np
sys.path.insert(0, ".")
from oracle import orb as oorb
from visual_odometry_ros_b200 import capi, synth

rng = np.random.default_rng(5)
ctx = capi.Context(device=0, max_w=1920, max_h=1200, n_slots=1, max_feat=8192)
ctx.set_detector("orb", 20)
bad = n = 0
for (w, h), bins in [((1241, 376), [(24, 12), (30, 12), (64, 32), (7, 3)]), ((640, 480), [(16, 12), (20, 15), (33, 17)]),
                     ((333, 247), [(7, 5), (11, 9)]), ((1920, 1200), [(48, 30)]), ((752, 480), [(24, 16), (25, 16)])]:
    img = synth.textured_image(np.random.default_rng(w), w, h)
    ctx.upload_image(0, img)
    for (bu, bv) in bins:
        for occ_n in (0, 50, 400):
            occ = np.stack([rng.uniform(-5, w + 5, occ_n), rng.uniform(-5, h + 5, occ_n)], 1).astype(np.float32)
            for thr in (20,):
                g = ctx.detect_bucketed(0, occ, bu, bv)
                o = oorb.detect_bucketed(img, occ, bu, bv, thr, backend="cv2")
                n += 1
                if not np.array_equal(g, o):
                    bad += 1
                    print("MISMATCH", (w, h), (bu, bv), occ_n, len(g), len(o))
print("cases", n, "bad", bad)
