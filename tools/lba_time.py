"""Wall time of vo_lba_solve (host buffers in and out, one sync) at the two window sizes of the benchmark."""
import sys
import time

sys.path.insert(0, ".")
from visual_odometry_ros_b200 import capi, synth  # noqa: E402

ctx = capi.Context(device=0, max_w=64, max_h=64, n_slots=0, max_feat=64)
for M in (5000, 7500):
    p = synth.lba_problem(seed=4004, n_kf=10, n_points=M)
    for _ in range(5):
        ctx.lba_solve(p)
    t0 = time.perf_counter()
    for _ in range(50):
        ctx.lba_solve(p)
    print(f"M={M} observations={p['n_obs']}: {(time.perf_counter() - t0) / 50 * 1e3:.3f} ms per call")
ctx.close()
