"""Pose-only GN (stereo / mono, both variants) and local BA against the C restatements over many random problems."""
import sys
import numpy as np
sys.path.insert(0, ".")
from oracle import lba as olba, pose as opose
from visual_odometry_ros_b200 import capi, synth

ctx = capi.Context(device=0, max_w=64, max_h=64, n_slots=0, max_feat=8192)
K, Tlr = synth.kitti_K(), synth.kitti_T_lr()
bad = n = 0
worst = 0.0
same_it = mask_eq = 0
worst_same = 0.0
for npts in (11, 37, 100, 500, 2000, 5000):
    for seed in range(6):
        for outl, noise in ((0.1, 0.3), (0.3, 0.5), (0.0, 0.05)):
            s = synth.pose_scene(seed=100 * seed + npts, n=npts, outlier_frac=outl, noise_px=noise,
                                 rotvec=(0.002 * seed, -0.012 + 0.004 * seed, 0.001), t=(0.02, -0.01 * seed, 0.85))
            ok_o, T_o, m_o, it_o = opose.pose_gn_stereo(s["X"], s["pts_l1"], s["pts_r1"], K, K, Tlr, 3.0, np.eye(4))
            ok_g, T_g, m_g, it_g = ctx.pose_gn_stereo(s["X"], s["pts_l1"], s["pts_r1"], K, K, Tlr, 3.0, np.eye(4))
            d = float(np.abs(T_g - T_o).max())
            same_it += it_g == it_o
            mask_eq += np.array_equal(m_g, m_o)
            if it_g == it_o:
                worst_same = max(worst_same, d)
            good = ok_g == ok_o and np.array_equal(m_g, m_o) and d <= 1e-6 and it_g == it_o
            for variant in (0, 1):
                r_o = opose.pose_gn_mono(s["X"], s["pts_l1"], K, 5, np.eye(3), np.zeros(3), variant)
                r_g = ctx.pose_gn_mono(s["X"], s["pts_l1"], K, 5, np.eye(3), np.zeros(3), variant)
                dm = max(float(np.abs(r_g[1] - r_o[1]).max()), float(np.abs(r_g[2] - r_o[2]).max()))
                good &= r_g[0] == r_o[0] and np.array_equal(r_g[3], r_o[3]) and dm <= 1e-6
                d = max(d, dm)
            worst = max(worst, d)
            n += 1
            bad += not good
            if not good and abs(it_g - it_o) > 1:
                print("POSE MISMATCH n", npts, "seed", seed, outl, "dT", d, "iters", it_g, it_o, "mask equal", np.array_equal(m_g, m_o))
print("pose cases", n, "beyond 1e-6 or different stop iteration", bad, "| stereo: same stop iteration", same_it, "identical masks", mask_eq,
      "| worst |dT| with the same iteration count %.2e, overall %.2e" % (worst_same, worst))
badl = nl = 0
worstl = 0.0
for (M, nkf, stereo) in [(1, 3, True), (7, 3, False), (50, 4, True), (300, 6, False), (1000, 9, True), (3000, 10, True), (800, 10, False)]:
    for seed in range(4):
        p = synth.lba_problem(seed=seed * 17 + M, n_kf=nkf, n_points=M, stereo=stereo)
        rc, poses_o, points_o, avg_o, ok_o = olba.lba_solve(p)
        poses_g, points_g, avg_g, ok_g = ctx.lba_solve(p)
        dp, dx, de = float(np.abs(poses_g - poses_o).max()), float(np.abs(points_g - points_o).max()), float(np.abs(avg_g - avg_o).max())
        # bars of tests/test_lba_gpu.py: poses 1e-6; landmark positions 1e-6 except near-singular C_i, where the oracle itself moves by
        # 2e-6 under a landmark permutation (bounded at 2e-5); M = 1 is rank-deficient (reported, not judged)
        good = (dp <= 1e-6 and dx <= 2e-5 and bool(ok_g) == bool(ok_o)) or M == 1
        worstl = max(worstl, dp if M > 1 else 0.0)
        nl += 1
        badl += not good
        if dx > 1e-6 or not good:
            print("LBA note" if good else "LBA MISMATCH", M, nkf, stereo, seed, "pose %.2e landmark %.2e avg_err %.2e" % (dp, dx, de))
print("lba cases", nl, "bad", badl, "worst pose deviation (M > 1) %.2e" % worstl)
