"""Teacher-forced stereo frame step against the oracle composition over several rendered sequences (different seeds),
both extractors: statistics of survivor-set agreement, gate counts, pixel and pose differences."""
import sys
import numpy as np
import torch
sys.path.insert(0, ".")
from oracle import stereo_vo as osvo
from visual_odometry_ros_b200 import capi, synth

W, H = synth.SMALL_W, synth.SMALL_H
K, Tlr = synth.small_K(), synth.kitti_T_lr()
NBU, NBV = 32, 12
tot = dict(frames=0, idx_equal=0, counts_equal=0, new_equal=0, agree=0, total=0, px_ok=0, px_tot=0)
worst_dt = worst_px = 0.0
for seed in (3103, 41, 77, 1234):
    L, R, T = synth.stereo_sequence(12, W, H, K, seed=seed, device="cuda")
    for det in ("harris", "orb"):
        prm = osvo.default_params(n_bins_u=NBU, n_bins_v=NBV, kf_trans=2.0, detector=det, fast_threshold=20)
        vo = osvo.StereoVOOracle(W, H, K, K, Tlr, prm)
        ctx = capi.Context(device=0, max_w=W, max_h=H, n_slots=4, max_feat=4096)
        ctx.set_detector(det, 20)
        common = dict(K_l=K, K_r=K, T_lr=Tlr, win=prm["window_size"], max_level=prm["max_level"], thres_err=prm["thres_error"],
                      thres_poseba=prm["thres_poseba_error"], thres_bi=prm["thres_bidirection"], n_bins_u=NBU, n_bins_v=NBV)
        for k in range(len(L)):
            vo.track(L[k], R[k])
            dbg = vo.dbg
            sl, sr, sp = 2 * (k % 2), 2 * (k % 2) + 1, 2 * ((k + 1) % 2)
            if k == 0:
                ctx.stereo_frame_step(-1, sl, sr, L[k], R[k], np.zeros((0, 2)), np.zeros((0, 2)), np.zeros((0, 3)), np.zeros(0), None, None,
                                      new_depth_gate=False, **common)
                continue
            g = ctx.stereo_frame_step(sp, sl, sr, L[k], R[k], dbg["pts_l0"], dbg["pts_r0"], dbg["Xw"], dbg["tri"], dbg["T_wp"], dbg["dT_prev"], **common)
            st = dbg["step"]
            tot["frames"] += 1
            tot["agree"] += len(np.intersect1d(g["index"], st["index"])); tot["total"] += max(len(g["index"]), len(st["index"]))
            if np.array_equal(g["index"], st["index"]):
                tot["idx_equal"] += 1
                tot["counts_equal"] += g["counts"] == st["counts"]
                d = np.concatenate([np.abs(g["pts_l1"] - st["pts_l1"]).max(1), np.abs(g["pts_r1"] - st["pts_r1"]).max(1)])
                tot["px_ok"] += int((d <= 0.01).sum()); tot["px_tot"] += len(d)
                worst_px = max(worst_px, float(d.max()))
                worst_dt = max(worst_dt, float(np.abs(g["dT_pc"] - st["dT_pc"]).max()))
                tot["new_equal"] += (len(g["new_l1"]) == len(dbg.get("new_l", [])) and np.array_equal(g["new_l1"], dbg.get("new_l", np.zeros((0, 2)))))
            else:
                print("  seed", seed, det, "frame", k, "survivor sets differ:", len(g["index"]), len(st["index"]),
                      "sym diff", len(np.setxor1d(g["index"], st["index"])))
        ctx.close()
print(tot, "worst |dT| %.2e" % worst_dt, "worst px %.4f" % worst_px)
