#!/usr/bin/env python
"""Run the StereoVO (or, with --mono, the MonoVO) drop-in over a rendered corridor sequence (BASELINE config 3) and print
per-frame timings.
Usage: python tools/run_sequence.py [--frames N] [--small] [--bins U V] [--mono] [--detector harris|orb]"""
import argparse
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--frames", type=int, default=30)
    ap.add_argument("--small", action="store_true")
    ap.add_argument("--bins", type=int, nargs=2, default=[64, 32])
    ap.add_argument("--mono", action="store_true", help="MonoVO::trackImage on the left images")
    ap.add_argument("--detector", default="harris", choices=["harris", "orb"], help="K-det or the reference's cv::ORB restated")
    args = ap.parse_args()
    import torch
    from visual_odometry_ros_b200 import stereo_vo as svo, synth
    w, h, K = (synth.SMALL_W, synth.SMALL_H, synth.small_K()) if args.small else (synth.KITTI_W, synth.KITTI_H, synth.kitti_K())
    dev = "cuda" if torch.cuda.is_available() else "cpu"
    L, R, T = synth.stereo_sequence(args.frames, w, h, K, seed=3003, device=dev)
    T0inv = np.linalg.inv(T[0])
    if args.mono:
        from visual_odometry_ros_b200 import mono_vo as mvo
        vo = mvo.MonoVO(mvo.make_parameters(w, h, K, max_level=3, n_bins_u=args.bins[0], n_bins_v=args.bins[1], detector=args.detector))
        scale = None
        for k in range(args.frames):
            t0 = time.perf_counter()
            vo.trackImage(L[k], 0.1 * k)
            ms = (time.perf_counter() - t0) * 1e3
            fi = vo.frame_info()
            gt = T0inv @ T[k]
            if k == 1:
                scale = np.linalg.norm(gt[:3, 3]) / max(1e-9, np.linalg.norm(vo.pose()[:3, 3]))     # |t| = 1 at initialisation
            err = np.abs((scale or 1.0) * vo.pose()[:3, 3] - gt[:3, 3]).max()
            print(f"frame {k:4d}  {ms:7.3f} ms  kf={fi['keyframe']}  in={fi['n_in']:5d} tracked={fi['n_tracked']:5d} new={fi['n_new']:4d} "
                  f"5pt={fi['used_5point']} recon={fi['n_recon']:4d} lba={fi['lba_points']:5d}/{fi['lba_obs']:6d}  |s t - t_gt|max={err:.4f} m  "
                  f"[step {fi['ms_step']:.2f} book {fi['ms_book']:.2f} recon {fi['ms_recon']:.2f} pack {fi['ms_lba_pack']:.2f} "
                  f"solve {fi['ms_lba_solve']:.2f} stats {fi['ms_stats']:.2f}]")
        vo.close()
        return
    vo = svo.StereoVO(svo.make_parameters(w, h, K, K, synth.kitti_T_lr(), n_bins_u=args.bins[0], n_bins_v=args.bins[1], detector=args.detector))
    for k in range(args.frames):
        t0 = time.perf_counter()
        vo.trackStereoImages(L[k], R[k], 0.1 * k)
        ms = (time.perf_counter() - t0) * 1e3
        fi = vo.frame_info()
        gt = T0inv @ T[k]
        err = np.abs(vo.pose()[:3, 3] - gt[:3, 3]).max()
        print(f"frame {k:4d}  {ms:7.3f} ms  kf={fi['keyframe']}  in={fi['n_in']:5d} tracked={fi['n_tracked']:5d} new={fi['n_new']:4d} "
              f"lba={fi['lba_points']:5d}/{fi['lba_obs']:6d}  |t - t_gt|max={err:.4f} m  "
              f"[step {fi['ms_step']:.2f} book {fi['ms_book']:.2f} recon {fi['ms_recon']:.2f} pack {fi['ms_lba_pack']:.2f} "
              f"solve {fi['ms_lba_solve']:.2f} stats {fi['ms_stats']:.2f}]")
    vo.close()


if __name__ == "__main__":
    main()
