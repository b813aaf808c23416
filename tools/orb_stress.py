"""K-orb vs cv2.ORB on assorted image sizes / contents / FAST thresholds (keypoint sets must be identical)."""
import sys
import numpy as np
import cv2
sys.path.insert(0, ".")
from visual_odometry_ros_b200 import capi, synth


def ref(img, thr):
    o = cv2.ORB_create()
    o.setMaxFeatures(10000); o.setScaleFactor(1.2); o.setNLevels(8); o.setEdgeThreshold(31); o.setFirstLevel(0); o.setWTA_K(2)
    o.setScoreType(cv2.ORB_HARRIS_SCORE); o.setPatchSize(31); o.setFastThreshold(thr)
    return sorted((k.octave, np.float32(k.pt[1]), np.float32(k.pt[0]), np.float32(k.response)) for k in o.detect(img, None))


rng = np.random.default_rng(123)
ctx = capi.Context(device=0, max_w=1920, max_h=1200, n_slots=1, max_feat=1024)
bad = 0
cases = []
for (w, h) in [(752, 480), (333, 247), (1920, 1200), (640, 480), (129, 97), (1241, 376)]:
    tex = synth.textured_image(np.random.default_rng(w), w, h)
    noise = rng.integers(0, 256, (h, w), dtype=np.uint8)
    blocks = (np.kron(rng.integers(0, 2, ((h + 15) // 16, (w + 15) // 16)), np.ones((16, 16)))[:h, :w] * 200 + 20).astype(np.uint8)
    smooth = cv2.GaussianBlur(noise, (0, 0), 3.0)
    for name, img in (("texture", tex), ("noise", noise), ("blocks", blocks), ("smooth", smooth)):
        for thr in (5, 20, 60):
            cases.append((w, h, name, np.ascontiguousarray(img), thr))
for w, h, name, img, thr in cases:
    ctx.upload_image(0, img)
    P, R, O = ctx.orb_detect(0, thr, max_keypoints=40000)
    got = sorted((int(o), np.float32(p[1]), np.float32(p[0]), np.float32(r)) for p, r, o in zip(P, R, O))
    r = ref(img, thr)
    ok = got == r
    bad += not ok
    if not ok:
        print("MISMATCH", w, h, name, thr, len(got), len(r), sorted(set(got) - set(r))[:3], sorted(set(r) - set(got))[:3])
print("cases", len(cases), "mismatches", bad)
