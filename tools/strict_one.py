"""One strict-order pose solve (N = 2000) for profiling."""
import sys
import numpy as np
sys.path.insert(0, ".")
from visual_odometry_ros_b200 import capi, synth
ctx = capi.Context(device=0, max_w=64, max_h=64, n_slots=0, max_feat=8192)
s = synth.pose_scene(seed=1001, n=2000)
K, Tlr = synth.kitti_K(), synth.kitti_T_lr()
for _ in range(4):
    r = ctx.pose_gn_stereo(s["X"], s["pts_l1"], s["pts_r1"], K, K, Tlr, 3.0, np.eye(4), flags=capi.VO_POSE_STRICT)
print("iters", r[3])
ctx.close()
