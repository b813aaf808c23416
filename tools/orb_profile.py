"""Profiling driver: K-orb (cv::ORB::detect restated) + bucketing on one KITTI-size image, three times."""
import sys
import time
import numpy as np
sys.path.insert(0, ".")
from visual_odometry_ros_b200 import capi, synth

L, _, _ = synth.stereo_sequence(2, synth.KITTI_W, synth.KITTI_H, synth.kitti_K(), seed=3003, device="cuda")
ctx = capi.Context(device=0, max_w=1241, max_h=376, n_slots=2, max_feat=4096)
ctx.upload_image(0, L[0])
ctx.set_detector("orb", 20)
occ = np.zeros((0, 2), np.float32)
for _ in range(3):
    pts = ctx.detect_bucketed(0, occ, 64, 32)
t0 = time.perf_counter()
for _ in range(50):
    pts = ctx.detect_bucketed(0, occ, 64, 32)
print("orb bucketed detect, host call incl. sync and D2H: %.3f ms, %d points" % ((time.perf_counter() - t0) * 20, len(pts)))
ctx.set_detector("harris")
for _ in range(3):
    ctx.detect_bucketed(0, occ, 64, 32)
t0 = time.perf_counter()
for _ in range(50):
    pts = ctx.detect_bucketed(0, occ, 64, 32)
print("K-det bucketed detect: %.3f ms, %d points" % ((time.perf_counter() - t0) * 20, len(pts)))
import cv2
from oracle import orb as oorb
cv2.setNumThreads(16)
oorb.detect_cv2(L[0], 20)
t0 = time.perf_counter()
for _ in range(10):
    oorb.detect_cv2(L[0], 20)
print("cv2.ORB.detect (16 threads available): %.2f ms" % ((time.perf_counter() - t0) * 100))
