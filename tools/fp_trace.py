"""Profiling aid: phase clocks of the five-point minimal solver (run with VO_5PT_TRACE=1)."""
import numpy as np, sys
sys.path.insert(0, ".")
from visual_odometry_ros_b200 import capi, synth
ctx = capi.Context(device=0, max_w=1241, max_h=376, n_slots=2, max_feat=4096)
rng = np.random.default_rng(5)
sets = []
for k in range(64):
    R = synth.so3_exp(rng.normal(0, 0.08, 3)); t = rng.normal(0, 1, 3); t /= np.linalg.norm(t)
    X = np.stack([rng.uniform(-6, 6, 5), rng.uniform(-3, 3, 5), rng.uniform(3, 40, 5)], 1)
    X1 = X @ R.T + t
    sets.append(np.stack([X[:, 0] / X[:, 2], X[:, 1] / X[:, 2], X1[:, 0] / X1[:, 2], X1[:, 1] / X1[:, 2]], 1))
for _ in range(3):
    ctx.five_point_minimal(np.asarray(sets))
