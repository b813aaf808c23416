"""C++ MonoVO next to the oracle composition (five-point hook = the CUDA stage with the same seed) over several rendered
sequences and both extractors: frames in step, keyframe / five-point / reconstruction / LBA agreement, pose differences."""
import sys
import numpy as np
sys.path.insert(0, ".")
from oracle import mono_vo as omvo
from visual_odometry_ros_b200 import capi, mono_vo as mvo, synth

W, H = synth.SMALL_W, synth.SMALL_H
K = synth.small_K()
NBU, NBV, SEED = 32, 12, 11
ctx = capi.Context(device=0, max_w=W, max_h=H, n_slots=2, max_feat=4096)
for seed in (3103, 41, 77):
    L, _, T = synth.stereo_sequence(16, W, H, K, seed=seed, device="cuda")
    for det in ("harris", "orb"):
        def fp_gpu(frame_id, p0, p1):
            r = ctx.pose_5point(p0, p1, K, 1.0, seed=SEED + frame_id)
            return True, r["R10"], r["t10"], r["mask"]
        ora = omvo.MonoVOOracle(W, H, K, omvo.default_params(n_bins_u=NBU, n_bins_v=NBV, max_level=3, kf_trans=2.0, detector=det, fast_threshold=15),
                                five_point=fp_gpu)
        vo = mvo.MonoVO(mvo.make_parameters(W, H, K, max_level=3, n_bins_u=NBU, n_bins_v=NBV, thres_translation=2.0, seed=SEED, detector=det,
                                            thres_fastscore=15))
        in_step, first_flip, kf_eq, rec_eq, lba_eq, worst = True, None, 0, 0, 0, 0.0
        for k in range(len(L)):
            Twc_o, info = ora.track(L[k])
            vo.trackImage(L[k], 0.1 * k)
            fi = vo.frame_info()
            ids, pts = vo.tracks()
            same = np.array_equal(ids, ora.prev.lm_ids)
            if in_step and not same:
                in_step, first_flip = False, k
            kf_eq += fi["keyframe"] == int(info["keyframe"])
            if in_step:
                rec_eq += fi["n_recon"] == info["n_recon"] + info.get("n_recon_kf", 0)
                lba_eq += (info["lba"] is None and fi["lba_points"] == 0) or (info["lba"] is not None and fi["lba_points"] == info["lba"]["n_points"])
                worst = max(worst, float(np.abs(vo.pose()[:3, 3] - Twc_o[:3, 3]).max()))
        n_step = first_flip if first_flip is not None else len(L)
        print(f"seed {seed} {det:6s}: in step for {n_step}/{len(L)} frames, keyframe decisions equal {kf_eq}/{len(L)}, reconstruction counts {rec_eq}/{n_step}, "
              f"LBA sizes {lba_eq}/{n_step}, worst |dt| while in step {worst:.2e} m, stats consistent {vo.stats_consistent()}")
        vo.close()
