"""Per-kernel times of the local BA (BASELINE config 4: 10 keyframes x 5000 landmarks) with VO_LBA_TRACE=1: build / reduce+solve per
LM iteration from CUDA events, phase clocks of k_lba_solve, and the host-call wall time."""
import os
import sys
import time

os.environ["VO_LBA_TRACE"] = "1"
sys.path.insert(0, ".")
from visual_odometry_ros_b200 import capi, synth  # noqa: E402

ctx = capi.Context(device=0, max_w=64, max_h=64, n_slots=0, max_feat=64)
for M in (5000, 7500):
    p = synth.lba_problem(seed=4004, n_kf=10, n_points=M)
    for _ in range(2):
        ctx.lba_solve(p)
    print(f"--- M={M} obs={p['n_obs']}", flush=True)
    ctx.lba_solve(p)
os.environ.pop("VO_LBA_TRACE")
ctx.close()
