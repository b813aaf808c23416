"""Timing of the pose-GN accumulation modes (VO_POSE_FAST: FP64 tree / VO_POSE_STRICT: sequential FP32 in point order):
one frame-step-sized problem through the host entry, and the batched config-1 workload (4096 x 500 points)."""
import json
import sys
import time

import numpy as np
import torch

sys.path.insert(0, ".")
from visual_odometry_ros_b200 import capi, synth  # noqa: E402

ctx = capi.Context(device=0, max_w=64, max_h=64, n_slots=0, max_feat=8192)
dev = torch.device("cuda:0")
K, Tlr = synth.kitti_K(), synth.kitti_T_lr()
out = {}
for n in (500, 2000, 2800):
    s = synth.pose_scene(seed=1001, n=n)
    for name, fl in (("fast", capi.VO_POSE_FAST), ("strict", capi.VO_POSE_STRICT)):
        for _ in range(5):
            r = ctx.pose_gn_stereo(s["X"], s["pts_l1"], s["pts_r1"], K, K, Tlr, 3.0, np.eye(4), flags=fl)
        t0 = time.perf_counter()
        reps = 200
        for _ in range(reps):
            r = ctx.pose_gn_stereo(s["X"], s["pts_l1"], s["pts_r1"], K, K, Tlr, 3.0, np.eye(4), flags=fl)
        ms = (time.perf_counter() - t0) / reps * 1e3
        out[f"single_n{n}_{name}"] = {"ms_host_call": ms, "iters": r[3]}
nprob, npts = 4096, 500
s = synth.pose_scene(seed=1001, n=npts)
rng = np.random.default_rng(5)
X = np.tile(s["X"][None], (nprob, 1, 1)).astype(np.float32)
pl = (s["pts_l1"][None] + rng.normal(0, 0.05, (nprob, npts, 2))).astype(np.float32)
pr = (s["pts_r1"][None] + rng.normal(0, 0.05, (nprob, npts, 2))).astype(np.float32)
X_d, pl_d, pr_d = (torch.from_numpy(a).to(dev) for a in (X, pl, pr))
off_d = torch.arange(0, (nprob + 1) * npts, npts, dtype=torch.int32, device=dev)
T_d = torch.eye(4, device=dev).repeat(nprob, 1, 1).contiguous()
eye = T_d.clone()
mask_d = torch.zeros(nprob * npts, dtype=torch.uint8, device=dev)
it_d = torch.zeros(nprob, dtype=torch.int32, device=dev)
ok_d = torch.zeros(nprob, dtype=torch.int32, device=dev)
stream = torch.cuda.Stream(device=dev)      # an explicit stream shared by torch and the context (handle 0 would mean "own stream")
torch.cuda.set_stream(stream)
ctx2 = capi.Context(device=0, max_w=64, max_h=64, n_slots=0, max_feat=64, stream=stream.cuda_stream)
res_T = {}
for name, fl in (("fast", capi.VO_POSE_FAST), ("strict", capi.VO_POSE_STRICT)):
    def solve():
        T_d.copy_(eye)
        ctx2.pose_gn_stereo_batch_d(nprob, off_d.data_ptr(), X_d.data_ptr(), pl_d.data_ptr(), pr_d.data_ptr(), K, K, Tlr, 3.0,
                                    T_d.data_ptr(), mask_d.data_ptr(), ok_d.data_ptr(), it_d.data_ptr(), flags=fl)
    for _ in range(3):
        solve()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record(stream)
    for _ in range(10):
        solve()
    b.record(stream)
    torch.cuda.synchronize()
    ms = a.elapsed_time(b) / 10
    res_T[name] = T_d.cpu().numpy().copy()
    out[f"batch_{name}"] = {"ms_per_batch": ms, "solves_per_s": nprob / (ms * 1e-3), "mean_iters": float(it_d.float().mean().item())}
out["batch_fast_vs_strict_max_dT"] = float(np.abs(res_T["fast"] - res_T["strict"]).max())
print(json.dumps(out, indent=1))
