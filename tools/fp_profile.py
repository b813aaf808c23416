"""Profiling driver: three vo_pose_5point calls at 2000 correspondences + 12 MonoVO frames at KITTI size."""
import sys
import numpy as np
import torch
sys.path.insert(0, ".")
from visual_odometry_ros_b200 import capi, synth
from visual_odometry_ros_b200 import mono_vo as mvo

ctx = capi.Context(device=0, max_w=1241, max_h=376, n_slots=2, max_feat=4096)
sc = synth.two_view_scene(seed=1, n=2000)
for _ in range(3):
    ctx.pose_5point(sc["pts0"], sc["pts1"], sc["K4"], 1.0)
ctx.close()
L, _, _ = synth.stereo_sequence(12, synth.KITTI_W, synth.KITTI_H, synth.kitti_K(), seed=3003, device="cuda")
vo = mvo.MonoVO(mvo.make_parameters(synth.KITTI_W, synth.KITTI_H, synth.kitti_K(), max_level=3, n_bins_u=64, n_bins_v=32))
for k in range(12):
    vo.trackImage(L[k], 0.1 * k)
print("ok", vo.frame_info())
