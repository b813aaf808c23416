"""Undistortion on the device against cv2.remap + convertTo directly (maps read back from the device) for assorted sizes
and distortion strengths."""
import sys
import numpy as np
import cv2
sys.path.insert(0, ".")
from oracle import rectify as orect
from visual_odometry_ros_b200 import capi, synth

bad = n = 0
for (w, h) in [(752, 480), (1241, 376), (640, 480), (333, 247), (1920, 1200)]:
    ctx = capi.Context(device=0, max_w=w, max_h=h, n_slots=1, max_feat=1024)
    K4 = np.array([0.6 * w, 0.61 * w, 0.49 * w, 0.52 * h], np.float32)
    img = synth.textured_image(np.random.default_rng(w), w, h)
    for D5 in ((0, 0, 0, 0, 0), (-0.28, 0.07, 2e-4, 2e-5, 0.0), (0.15, -0.3, -1e-3, 2e-3, 0.1), (-0.6, 0.4, 0, 0, -0.1)):
        D5 = np.asarray(D5, np.float32)
        ctx.undistort_init(K4, D5, w, h)
        mu, mv = ctx.read_rectify_maps(0)
        ou, ov = orect.undistort_maps(K4, D5, w, h)
        ctx.upload_image_rectified(0, 0, img)
        ctx.build_pyramids(np.array([0], np.int32), 1, True)
        got = ctx.read_pyramid_level(0, 0)[0]
        ref = cv2.remap(img.astype(np.float32), mu, mv, cv2.INTER_LINEAR)
        ref8 = np.clip(np.rint(ref), 0, 255).astype(np.uint8)          # convertTo(CV_8UC1): saturate_cast<uchar>(cvRound)
        ok = np.array_equal(mu, ou) and np.array_equal(mv, ov) and np.array_equal(got, ref8)
        n += 1
        bad += not ok
        if not ok:
            print("MISMATCH", (w, h), D5, "maps", np.array_equal(mu, ou), np.array_equal(mv, ov), "image diff px", int((got != ref8).sum()))
    ctx.close()
print("cases", n, "bad", bad)
