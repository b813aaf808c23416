"""Smallest end-to-end run of the TMA-staged LK kernel (one pair, 300 features) against cv2; exits non-zero on any CUDA error.
Used as the first thing on a GPU box and under compute-sanitizer when something faults."""
import sys

import numpy as np

sys.path.insert(0, ".")
from visual_odometry_ros_b200 import capi, synth  # noqa: E402
from oracle import klt as oklt  # noqa: E402

win = int(sys.argv[1]) if len(sys.argv) > 1 else 21
ctx = capi.Context(device=0, max_w=640, max_h=192, n_slots=2, max_feat=1024)
rng = np.random.default_rng(7)
img0 = synth.textured_image(rng, 640, 192)
img1 = synth.warp_translate_field(img0, 2.3, -1.1)
pts0 = synth.grid_features(rng, 300, 640, 192, nx=25, ny=12)
ctx.upload_image(0, img0)
ctx.upload_image(1, img1)
p_g, s_g, e_g = ctx.klt_track(0, 1, pts0, win, 3)
ctx.synchronize()
p_c, s_c, e_c = oklt.lk_cv2(img0, img1, pts0, win, 3)
ok = (s_g > 0) & (s_c > 0)
d = np.abs(p_g - p_c).max(1)[ok]
print(f"klt_smoke win={win}: status agreement {np.mean(s_g == s_c):.4f}, tracked {ok.sum()}, max|dp| {d.max():.3e} px, launches {ctx.launch_count}")
ctx.close()
sys.exit(0 if (np.mean(s_g == s_c) >= 0.999 and np.percentile(d, 99.9) <= 0.01) else 1)
