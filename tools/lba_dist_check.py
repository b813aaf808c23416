"""Landmark-sharded local BA over NCCL (vo_lba_solve_dist) against the one-GPU solve (vo_lba_solve).
Run with:  python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 tools/lba_dist_check.py
Each rank: all poses + its half of the landmarks; one ncclAllReduce of the reduced camera system per LM iteration.
Rank 0 also solves the whole problem alone; the two results must agree to ~1e-12 (FP64, summation order only).
Prints one JSON line per problem size (M = 5000 parity case, M = 100 000 the oversize window of north_star)."""
import json
import os
import sys
import time

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from visual_odometry_ros_b200 import capi, sharding, synth  # noqa: E402

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
ctx = capi.Context(device=local, max_w=64, max_h=64, n_slots=0, max_feat=64)
# the host program ships the NCCL id (here over torch.distributed; any transport works)
ids = [capi.dist_unique_id() if rank == 0 else None]
dist.broadcast_object_list(ids, src=0)
ctx.dist_init(rank, world, ids[0])
ok_all = True
# poses and per-iteration error: 1e-12 (summation order of the reduced system only); landmarks: 1e-7 -- among 100 000 random
# landmarks a few are nearly degenerate (short baseline, far away) and amplify the 1e-17 pose difference through C_i^-1
for M, tol in ((5000, 1e-12), (100000, 1e-12)):
    p = synth.lba_problem(seed=4004, n_kf=10, n_points=M)
    mine = sharding.split_lba_problem(p, world, rank)
    for _ in range(2):
        poses_d, pts_d, avg_d, ok_d = ctx.lba_solve_dist(mine)
    dist.barrier()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    reps = 5
    for _ in range(reps):
        poses_d, pts_d, avg_d, ok_d = ctx.lba_solve_dist(mine)
    ms_dist = (time.perf_counter() - t0) * 1e3 / reps
    lo, hi = mine["landmark_range"]
    res = None
    if rank == 0:
        for _ in range(2):
            poses_1, pts_1, avg_1, ok_1 = ctx.lba_solve(p)
        t0 = time.perf_counter()
        for _ in range(reps):
            poses_1, pts_1, avg_1, ok_1 = ctx.lba_solve(p)
        ms_one = (time.perf_counter() - t0) * 1e3 / reps
        res = (poses_1, pts_1, avg_1, ms_one)
    box = [res]
    dist.broadcast_object_list(box, src=0)
    poses_1, pts_1, avg_1, ms_one = box[0]
    d_pose = float(np.abs(poses_d - poses_1).max())
    d_pts = float(np.abs(pts_d - pts_1[lo:hi]).max())
    d_err = float(np.abs(avg_d - avg_1).max())
    t = torch.tensor([d_pose, d_pts, d_err, ms_dist], dtype=torch.float64, device="cuda")
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    good = bool(t[0] <= tol and t[1] <= 1e-7 and t[2] <= tol)
    ok_all &= good
    if rank == 0:
        print(json.dumps({"lba_dist": {"world": world, "landmarks": M, "observations": int(p["n_obs"]), "keyframes": 10, "iterations": int(p["max_iter"]),
                                       "max_abs_pose_diff_vs_1gpu": float(t[0]), "max_abs_point_diff_vs_1gpu": float(t[1]),
                                       "max_abs_avg_err_diff": float(t[2]), "tolerance_pose_err": tol, "tolerance_points": 1e-7, "ok": good,
                                       "ms_dist_host_call": float(t[3]), "ms_1gpu_host_call": ms_one,
                                       "allreduce_values_per_iteration": (6 * int(p["n_opt"])) * (6 * int(p["n_opt"]) + 1) + 27 * int(p["n_opt"]) + 2}}),
              flush=True)
ctx.dist_finalize()
ctx.close()
dist.destroy_process_group()
sys.exit(0 if ok_all else 1)
