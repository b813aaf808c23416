"""Five-point RANSAC against the ground truth for assorted motions and scenes (incl. a planar scene), beside cv2."""
import sys
import numpy as np
import cv2
sys.path.insert(0, ".")
from oracle import five_point as ofp
from visual_odometry_ros_b200 import capi, synth

ctx = capi.Context(device=0, max_w=1241, max_h=376, n_slots=1, max_feat=8192)


def ang(Ra, Rb):
    return float(np.arccos(np.clip((np.trace(np.asarray(Ra, np.float64) @ np.asarray(Rb, np.float64).T) - 1) / 2, -1, 1)))


def tdir(a, b):
    return float(np.degrees(np.arccos(np.clip(float(np.asarray(a, np.float64) @ np.asarray(b, np.float64)), -1, 1))))


bad = 0
cases = [("forward", (0.004, -0.02, 0.003), (0.05, -0.02, 0.9)), ("sideways", (0.0, 0.01, 0.0), (0.8, 0.0, 0.05)),
         ("vertical", (0.01, 0.0, 0.0), (0.0, 0.5, 0.1)), ("yaw 10 deg", (0.0, 0.17, 0.0), (0.1, 0.0, 0.9)),
         ("roll 5 deg", (0.0, 0.0, 0.087), (0.0, 0.0, 1.0)), ("backward", (0.0, 0.005, 0.0), (0.0, 0.0, -0.7))]
for name, rv, t in cases:
    for outl in (0.1, 0.4):
        for planar in (False, True):
            sc = synth.two_view_scene(seed=11, n=1200, rotvec=rv, t=t, outlier_frac=outl)
            if planar:      # all points on one slanted plane: a classical degeneracy for 8-point, fine for 5-point
                rng = np.random.default_rng(3)
                n = 1200
                X0 = np.stack([rng.uniform(-12, 12, n), rng.uniform(-3, 2, n), np.zeros(n)], 1)
                X0[:, 2] = 20 + 0.4 * X0[:, 0] - 0.8 * X0[:, 1]
                R10, t10 = sc["R10"], -sc["R10"] @ np.asarray(t, np.float64)
                X1 = X0 @ R10.T + t10
                K = sc["K4"]
                p0 = np.stack([K[0] * X0[:, 0] / X0[:, 2] + K[2], K[1] * X0[:, 1] / X0[:, 2] + K[3]], 1) + rng.normal(0, 0.3, (n, 2))
                p1 = np.stack([K[0] * X1[:, 0] / X1[:, 2] + K[2], K[1] * X1[:, 1] / X1[:, 2] + K[3]], 1) + rng.normal(0, 0.3, (n, 2))
                idx = rng.choice(n, int(outl * n), replace=False)
                p1[idx] += rng.choice([-1.0, 1.0], (len(idx), 2)) * rng.uniform(4, 40, (len(idx), 2))
                sc = dict(sc, pts0=p0.astype(np.float32), pts1=p1.astype(np.float32), t10=t10 / np.linalg.norm(t10))
            g = ctx.pose_5point(sc["pts0"], sc["pts1"], sc["K4"], 1.0, seed=1)
            ok, R_o, t_o, _, m_o, _ = ofp.calc_pose_5point(sc["pts0"], sc["pts1"], sc["K4"], 1.0)
            eg, ec = ang(g["R10"], sc["R10"]) * 1e3, ang(R_o, sc["R10"]) * 1e3
            dg, dc = tdir(g["t10"], sc["t10"]), tdir(t_o, sc["t10"])
            flag = "" if (eg < max(1.5 * ec, 3.0) and dg < max(1.5 * dc, 2.0)) else "  <-- worse than cv2"
            bad += bool(flag)
            print(f"{name:11s} outl {outl} planar {int(planar)}: rot err gpu {eg:6.2f} cv2 {ec:6.2f} mrad | t dir gpu {dg:5.2f} cv2 {dc:5.2f} deg | "
                  f"inliers gpu {int(g['mask'].sum())} cv2 {int(m_o.sum())}{flag}")
print("worse than cv2:", bad)
