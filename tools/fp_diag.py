import numpy as np, sys
sys.path.insert(0, ".")
from oracle import five_point as ofp
from visual_odometry_ros_b200 import capi, synth
ctx = capi.Context(device=0, max_w=1241, max_h=376, n_slots=2, max_feat=4096)
rng = np.random.default_rng(5)
sets = []
for k in range(200):
    R = synth.so3_exp(rng.normal(0, 0.08, 3)); t = rng.normal(0, 1, 3); t /= np.linalg.norm(t)
    X = np.stack([rng.uniform(-6, 6, 5), rng.uniform(-3, 3, 5), rng.uniform(3, 40, 5)], 1)
    X1 = X @ R.T + t
    q = np.stack([X[:, 0] / X[:, 2], X[:, 1] / X[:, 2], X1[:, 0] / X1[:, 2], X1[:, 1] / X1[:, 2]], 1)
    if k % 2: q += rng.normal(0, 1e-3, q.shape)
    sets.append(q)
got = ctx.five_point_minimal(np.asarray(sets))
worst = []
nsame = 0
for i, (q, G) in enumerate(zip(sets, got)):
    O = ofp.minimal_solutions(q)
    for g in G:
        worst.append((np.abs(ofp.cv_error(g, q)).max(), abs(np.linalg.det(g)), np.abs(2 * g @ g.T @ g - np.trace(g @ g.T) * g).max()))
    d = [min(min(np.abs(a - b).max(), np.abs(a + b).max()) for b in G) if len(G) else 9 for a in O]
    d2 = [min(min(np.abs(a - b).max(), np.abs(a + b).max()) for b in O) if len(O) else 9 for a in G]
    same = len(G) == len(O) and max(d + d2 + [0]) < 1e-6
    nsame += same
    if not same: print(i, len(G), len(O), ["%.1e" % v for v in d], ["%.1e" % v for v in d2])
w = np.asarray(worst)
print("n_same", nsame, "worst", w.max(0), "q99", np.quantile(w, 0.99, axis=0))
for seed, outl in [(6006, 0.25), (6007, 0.4), (6008, 0.1)]:
    sc = synth.two_view_scene(seed=seed, outlier_frac=outl)
    ok, R_o, t_o, X0_o, m_o, E_o = ofp.calc_pose_5point(sc["pts0"], sc["pts1"], sc["K4"], 1.0)
    def ang(Ra, Rb): return np.arccos(np.clip((np.trace(Ra.astype(np.float64) @ Rb.astype(np.float64).T) - 1) / 2, -1, 1))
    for s in (1, 2, 3):
        g = ctx.pose_5point(sc["pts0"], sc["pts1"], sc["K4"], 1.0, seed=s)
        print(seed, s, "rot gpu-cv %.2f mrad  gpu-gt %.2f  cv-gt %.2f | tdir gpu-cv %.2f deg gpu-gt %.2f cv-gt %.2f | agree %.4f inl gpu %d cv %d ransac %d" % (
            ang(g["R10"], R_o) * 1e3, ang(g["R10"], sc["R10"]) * 1e3, ang(R_o, sc["R10"]) * 1e3,
            np.degrees(np.arccos(np.clip(g["t10"] @ t_o, -1, 1))), np.degrees(np.arccos(np.clip(g["t10"] @ sc["t10"], -1, 1))),
            np.degrees(np.arccos(np.clip(t_o @ sc["t10"], -1, 1))), (g["mask"] == m_o).mean(), g["mask"].sum(), m_o.sum(), g["n_ransac"]))
import time
sc = synth.two_view_scene(seed=1, n=2000)
for H in (256, 1024, 4096):
    ctx.pose_5point(sc["pts0"], sc["pts1"], sc["K4"], 1.0, n_hypotheses=H)
    t0 = time.perf_counter()
    for _ in range(20): ctx.pose_5point(sc["pts0"], sc["pts1"], sc["K4"], 1.0, n_hypotheses=H)
    print("H", H, "ms/call", (time.perf_counter() - t0) * 50)

import cv2
cv2.setNumThreads(16)
Kcv = np.array([[sc["K4"][0], 0, sc["K4"][2]], [0, sc["K4"][1], sc["K4"][3]], [0, 0, 1]], np.float64)
for frac in (0.1, 0.25, 0.4):
    s2 = synth.two_view_scene(seed=2, n=2000, outlier_frac=frac)
    cv2.findEssentialMat(s2["pts0"], s2["pts1"], Kcv, cv2.RANSAC, 0.999, 1.0)
    t0 = time.perf_counter()
    for _ in range(10): cv2.findEssentialMat(s2["pts0"], s2["pts1"], Kcv, cv2.RANSAC, 0.999, 1.0)
    tc = (time.perf_counter() - t0) * 100
    ctx.pose_5point(s2["pts0"], s2["pts1"], s2["K4"], 1.0)
    t0 = time.perf_counter()
    for _ in range(10): ctx.pose_5point(s2["pts0"], s2["pts1"], s2["K4"], 1.0)
    print("outliers", frac, "cv2.findEssentialMat ms", tc, "vo_pose_5point ms", (time.perf_counter() - t0) * 100)
