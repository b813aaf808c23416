#!/usr/bin/env python
"""bench.py -- KLT features/s (+ pose-GN solves/s, ms/frame) at 1241x376 stereo on B200.

Contract: `python bench.py --gpus N --steps K --warmup W` prints ONE JSON line on rank 0.

Workload (BASELINE.json configs[1], batched as configs[4] prescribes): P independent
KITTI-size stereo pairs per GPU and step; one step == FeatureTracker::track on every pair ==
(4-level pyramid of both images + Scharr derivative pyramid of the first + pyramidal LK of
2000 features, 21x21 window), i.e. exactly what one cv::calcOpticalFlowPyrLK call does in the
reference (core/visual_odometry/feature_tracker.cpp:29).  `value` = features tracked per second,
whole job, with the images already resident in HBM (slot level 0); `e2e` = the same through the
host-buffer C-ABI call vo_ft_track_batch (pinned host images + points in, points + masks out,
all copies inside the timed region).

`--impl reference` times the reference's own CPU implementation of the same step
(cv2.calcOpticalFlowPyrLK 4.13 -- the library call the reference makes -- with all host threads).
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

W, H, NFEAT, WIN, MAXLVL, THRES_ERR = 1241, 376, 2000, 21, 3, 80.0
METRIC = "KLT features/s at 1241x376 stereo (2000 features, 4-level pyramid, 21x21 window)"


def load_peak():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    try:
        with open(p) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region."""

    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.lines, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if not self.proc:
            return None
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.lines:
            parts = [p.strip() for p in ln.split(",")]
            if len(parts) < 6:
                continue
            try:
                sm.append(float(parts[0]))
                mx.append(float(parts[1]))
            except ValueError:
                continue
            for nm, v in zip(names, parts[2:6]):
                if v.lower().startswith("active"):
                    reasons.add(nm)
        if not sm:
            return None
        return {"sm_mhz": float(np.median(sm)), "sm_max_mhz": float(np.max(mx)), "reasons": sorted(reasons),
                "samples": len(sm)}


def make_pairs(P, seed):
    """P distinct synthetic stereo pairs + feature sets (cheap variations of a few base pairs)."""
    from visual_odometry_ros_b200 import synth
    nbase = min(P, 4)
    bases = [synth.klt_stereo_case(seed=seed + 17 * k, n=NFEAT) for k in range(nbase)]
    rng = np.random.default_rng(seed)
    lefts, rights, pts = [], [], []
    for i in range(P):
        b = bases[i % nbase]
        sh = int(rng.integers(0, W)) if i >= nbase else 0
        lefts.append(np.ascontiguousarray(np.roll(b["left"], sh, axis=1)))
        rights.append(np.ascontiguousarray(np.roll(b["right"], sh, axis=1)))
        pts.append(synth.grid_features(rng, NFEAT, W, H))
    return lefts, rights, np.stack(pts).astype(np.float32)


def cpu_reference_rate(left, right, pts0, threads, budget_s):
    """cv2.calcOpticalFlowPyrLK (the reference's library call) on one pair, repeated."""
    import cv2
    from oracle import klt as oklt
    cv2.setNumThreads(threads)
    oklt.track(oklt.lk_cv2, left, right, pts0, WIN, MAXLVL, THRES_ERR)  # warm-up
    t0 = time.perf_counter()
    reps = 0
    while True:
        oklt.track(oklt.lk_cv2, left, right, pts0, WIN, MAXLVL, THRES_ERR)
        reps += 1
        dt = time.perf_counter() - t0
        if dt >= budget_s or reps >= 2000:
            break
    return reps * len(pts0) / dt, reps, dt


def workload_string(P):
    return (f"cfg2 stereo KLT 1241x376, 2000 features, 4 levels, win 21; {P} independent pairs "
            f"per GPU per step (cfg5 batching); step = pyramids(both) + Scharr + LK")


def l2_policy_string(P):
    """Timing rule: inputs (or the per-step working set) larger than the GPU's 126 MB L2 -- a property of the workload."""
    return ("inputs larger than L2" if 2 * P * W * H > 126e6 else "working set (pyramids+derivatives) larger than L2")


def config_dict(P):
    return {"workload": workload_string(P), "pairs_per_gpu": P, "features_per_pair": NFEAT, "window": WIN, "levels": MAXLVL + 1,
            "l2_policy": l2_policy_string(P)}


_REF = {}


def _ref_worker_init(P, seed):
    """One worker process = one host core running the reference's library call single-threaded on its share of the pairs."""
    import cv2
    cv2.setNumThreads(1)
    lefts, rights, pts = make_pairs(P, seed)
    _REF.update(lefts=lefts, rights=rights, pts=pts)


def _ref_worker_track(idx):
    from oracle import klt as oklt
    n_ok = 0
    for i in idx:
        p1, m = oklt.track(oklt.lk_cv2, _REF["lefts"][i], _REF["rights"][i], _REF["pts"][i], WIN, MAXLVL, THRES_ERR)
        n_ok += int(m.sum())
    return n_ok


def run_reference(args):
    """The reference's own CPU implementation of the SAME step the GPU arm runs: FeatureTracker::track on P independent stereo
    pairs (cv2.calcOpticalFlowPyrLK 4.13 -- the library call of feature_tracker.cpp:29 -- + the post-filter), two ways:
    (a) one pair per host core (a pool of single-threaded processes: what BASELINE.md section 3 promises for the batched
    config), (b) pair after pair with OpenCV's internal parallel_for on all threads.  value = the faster of the two."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import multiprocessing as mp
    P = args.pairs
    cores = os.cpu_count() or 1
    try:
        cores = len(os.sched_getaffinity(0))
    except Exception:
        pass
    # (a) process pool, created before this process touches cv2's thread pool
    ctxmp = mp.get_context("spawn")
    nproc = max(1, min(cores, P))
    shares = [list(range(k, P, nproc)) for k in range(nproc)]
    pool_val = pool_ms = None
    try:
        with ctxmp.Pool(nproc, initializer=_ref_worker_init, initargs=(P, 2002)) as pool:
            for _ in range(max(1, min(args.warmup, 2))):
                pool.map(_ref_worker_track, shares, chunksize=1)
            t0 = time.perf_counter()
            for _ in range(args.steps):
                pool.map(_ref_worker_track, shares, chunksize=1)
            dt_pool = time.perf_counter() - t0
        pool_val = args.steps * P * NFEAT / dt_pool
        pool_ms = 1e3 * dt_pool / args.steps
    except Exception as e:          # a box that cannot spawn processes still reports (b)
        pool_err = repr(e)
    # (b) OpenCV's own threading, pair after pair: a bounded sample of the step (8 pairs), scaled to P pairs
    import cv2
    from oracle import klt as oklt
    lefts, rights, pts = make_pairs(min(P, 8), 2002)
    cv2.setNumThreads(cores)
    nb = len(lefts)
    for i in range(nb):
        oklt.track(oklt.lk_cv2, lefts[i], rights[i], pts[i], WIN, MAXLVL, THRES_ERR)
    reps = max(1, min(args.steps, 5))
    t0 = time.perf_counter()
    for _ in range(reps):
        for i in range(nb):
            oklt.track(oklt.lk_cv2, lefts[i], rights[i], pts[i], WIN, MAXLVL, THRES_ERR)
    dt_thr = time.perf_counter() - t0
    thr_val = reps * nb * NFEAT / dt_thr
    if pool_val is not None and pool_val >= thr_val:
        val, ms_step, how = pool_val, pool_ms, f"{nproc} single-threaded processes, one pair per core at a time"
    else:
        val, ms_step, how = thr_val, 1e3 * P * NFEAT / thr_val, f"pair after pair, OpenCV parallel_for on {cores} threads (sample of {nb} pairs scaled to {P})"
    sample = (f"{args.steps} steps x {P} stereo pairs x {NFEAT} features: cv2.calcOpticalFlowPyrLK 4.13 + track() post-filter; "
              f"pool of {nproc} processes: {pool_val and round(pool_val)} features/s, OpenCV-threaded ({cores} threads): {round(thr_val)} features/s; "
              f"reported = the faster ({how})")
    print(json.dumps({
        "impl": "reference", "metric": METRIC, "value": val, "unit": "features/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_step, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "u8/int16 fixed-point + f32", "data": "synthetic",
        "config": config_dict(P),
        "cpu_baseline": {"value": val, "unit": "features/s", "cores": cores, "kind": "reference", "sample": sample,
                         "pool_value": pool_val, "opencv_threads_value": thr_val},
        "e2e": {"value": val, "unit": "features/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--pairs", type=int, default=128, help="independent stereo pairs per GPU per step")
    ap.add_argument("--cpu-budget", type=float, default=12.0, help="seconds of CPU baseline work (rank 0, N=1)")
    ap.add_argument("--skip-extras", action="store_true", help="skip pose-GN / single-frame / sequence side measurements")
    ap.add_argument("--sequences", type=int, default=64, help="config 5: independent stereo sequences, sharded over the ranks")
    ap.add_argument("--seq-frames", type=int, default=100, help="config 5: frames per sequence")
    ap.add_argument("--cfg3-frames", type=int, default=1000, help="config 3: frames of the single-sequence measurement")
    args = ap.parse_args()
    if args.warmup < 3 and args.impl == "ours":
        args.warmup = 3
    if args.impl == "reference":
        return run_reference(args)

    import torch
    import torch.distributed as dist
    from visual_odometry_ros_b200 import capi, synth

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the product path has no CPU fallback")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    numa_binding = bind_to_gpu_numa(local)      # before any pinned allocation (first touch)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    def barrier():
        if world > 1:
            dist.barrier()

    def max_over_ranks(x):
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    P, n, K, Wm = args.pairs, NFEAT, args.steps, args.warmup
    lefts, rights, pts0_h = make_pairs(P, 2002 + 1000 * rank)
    # an explicit non-default stream shared by torch (events, tensors) and the context (kernels):
    # the legacy default stream's handle is 0, which vo_ctx_create reads as "own stream"
    stream = torch.cuda.Stream(device=dev)
    torch.cuda.set_stream(stream)
    assert stream.cuda_stream != 0
    ctx = capi.Context(device=local, max_w=W, max_h=H, n_slots=2 * P, max_feat=P * n, stream=stream.cuda_stream)
    slots0 = np.arange(P, dtype=np.int32)
    slots1 = np.arange(P, 2 * P, dtype=np.int32)
    all_slots = np.arange(2 * P, dtype=np.int32)

    # pinned host copies (e2e path) and HBM-resident slots (value path)
    host_imgs = torch.empty((2 * P, H, W), dtype=torch.uint8).pin_memory()
    for i in range(P):
        host_imgs[i] = torch.from_numpy(lefts[i])
        host_imgs[P + i] = torch.from_numpy(rights[i])
    host_np = host_imgs.numpy()
    for s in range(2 * P):
        ctx.upload_image(s, host_np[s])
    ctx.synchronize()

    pts0_d = torch.from_numpy(pts0_h).to(dev)
    pts1_d = torch.zeros_like(pts0_d)
    status_d = torch.zeros((P, n), dtype=torch.uint8, device=dev)
    err_d = torch.zeros((P, n), dtype=torch.float32, device=dev)
    counters_d = torch.zeros(2 * capi.VO_MAX_LEVELS, dtype=torch.int64, device=dev)
    nlev = MAXLVL + 1

    def step(ev=None, counters=None):
        ctx.invalidate_pyramids(all_slots)
        if ev:
            ev[0].record(stream)
        ctx.build_pyramids(slots0, nlev, True)
        ctx.build_pyramids(slots1, nlev, False)
        if ev:
            ev[1].record(stream)
        ctx.klt_track_batch_d(slots0, slots1, pts0_d.data_ptr(), n, WIN, MAXLVL, 0, pts1_d.data_ptr(),
                              status_d.data_ptr(), err_d.data_ptr(), counters.data_ptr() if counters is not None else None)
        if ev:
            ev[2].record(stream)

    # algorithmic bytes of one step's LK work, from the kernel's own counters (untimed step)
    step(counters=counters_d)
    torch.cuda.synchronize()
    cnt = counters_d.cpu().numpy().reshape(-1, 2)
    n_templ, n_iter = int(cnt[:, 0].sum()), int(cnt[:, 1].sum())
    klt_bytes_per_step = (WIN + 1) ** 2 * (5 * n_templ + n_iter)
    tracked = int(status_d.sum().item())

    for _ in range(Wm):
        step()
    torch.cuda.synchronize()
    evs = [[torch.cuda.Event(enable_timing=True) for _ in range(3)] for _ in range(K)]
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    clocks = ClockSampler(local)
    launches0 = ctx.launch_count
    barrier()
    torch.cuda.synchronize()
    if rank == 0:
        clocks.start()
    e0.record(stream)
    for k in range(K):
        step(ev=evs[k])
    e1.record(stream)
    torch.cuda.synchronize()
    barrier()
    launches = ctx.launch_count - launches0
    ms_total = max_over_ranks(e0.elapsed_time(e1))
    clk = clocks.stop() if rank == 0 else None
    ms_step = ms_total / K
    value = world * P * n / (ms_step * 1e-3)
    pyr_ms = float(np.mean([e[0].elapsed_time(e[1]) for e in evs]))
    klt_ms = float(np.mean([e[1].elapsed_time(e[2]) for e in evs]))
    klt_launches_per_step = (P + 63) // 64
    peak, peak_src = load_peak()
    achieved = klt_bytes_per_step / (klt_ms * 1e-3) / 1e9

    # ---------------------------------------------------------------- e2e (host buffers through the C ABI)
    ptrs0 = [host_np[i].ctypes.data for i in range(P)]
    ptrs1 = [host_np[P + i].ctypes.data for i in range(P)]
    Ke = max(3, min(K, 10))

    # the caller's point / mask arrays live in page-locked memory like the images (a real feeder's ring buffers would)
    pts0_pin = torch.from_numpy(pts0_h).pin_memory().numpy()
    pt_out_pin = torch.zeros((P, n, 2), dtype=torch.float32).pin_memory().numpy()
    mask_pin = torch.ones((P, n), dtype=torch.uint8).pin_memory().numpy()

    def e2e_step():
        mask_pin.fill(1)                     # FeatureTracker::track starts from an all-true mask
        return ctx.ft_track_batch(slots0, slots1, ptrs0, ptrs1, W, H, W, pts0_pin, WIN, MAXLVL, THRES_ERR, pts_track=pt_out_pin, mask=mask_pin)

    for _ in range(2):
        e2e_step()
    barrier()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    e0.record(stream)
    for _ in range(Ke):
        pt_e, m_e = e2e_step()
    e1.record(stream)
    torch.cuda.synchronize()
    wall = (time.perf_counter() - t0) * 1e3
    barrier()
    e2e_ms = max_over_ranks(max(e0.elapsed_time(e1), wall)) / Ke
    e2e_val = world * P * n / (e2e_ms * 1e-3)

    # anatomy of the e2e step: the same call with the images already in their slots (pyramids invalidated, so everything but
    # the image DMA still happens: point upload, the ramped chunks on two streams, result download, the sync)
    def e2e_resident_step():
        mask_pin.fill(1)
        ctx.invalidate_pyramids(all_slots)
        return ctx.ft_track_batch(slots0, slots1, None, None, W, H, W, pts0_pin, WIN, MAXLVL, THRES_ERR, pts_track=pt_out_pin, mask=mask_pin)
    for _ in range(2):
        e2e_resident_step()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(Ke):
        e2e_resident_step()
    torch.cuda.synchronize()
    e2e_resident_ms = (time.perf_counter() - t0) * 1e3 / Ke
    h2d = 2 * P * W * H + P * n * 8 + P * n
    d2h = P * n * 8 + P * n

    # ---------------------------------------------------------------- H2D-only bandwidth, all ranks copying at once
    # (what bounds e2e at N > 1: every rank's pinned images cross the same host memory / PCIe root complex)
    raw_d = torch.empty((2 * P, H, W), dtype=torch.uint8, device=dev)
    for _ in range(2):
        raw_d.copy_(host_imgs, non_blocking=True)
    barrier()
    torch.cuda.synchronize()
    e0.record(stream)
    for _ in range(5):
        raw_d.copy_(host_imgs, non_blocking=True)
    e1.record(stream)
    torch.cuda.synchronize()
    h2d_gbs_rank = 5 * host_imgs.numel() / (e0.elapsed_time(e1) * 1e-3) / 1e9
    if world > 1:
        t = torch.tensor([h2d_gbs_rank], dtype=torch.float64, device=dev)
        g = [torch.zeros_like(t) for _ in range(world)]
        dist.all_gather(g, t)
        h2d_gbs_all = [float(x.item()) for x in g]
    else:
        h2d_gbs_all = [h2d_gbs_rank]
    del raw_d
    # e2e cannot beat max(device time, image DMA time): the ceiling the measured H2D rate allows
    e2e_ceiling = world * P * n / (max(ms_step, 2 * P * W * H / (min(h2d_gbs_all) * 1e9) * 1e3) * 1e-3)
    numa = gpu_numa_info(local)

    # ---------------------------------------------------------------- BASELINE config 5 at SEQUENCE level: 64 independent
    # StereoVO sequences sharded 64 / N per rank (sharding.shard_range), no inter-GPU traffic
    seqs = None
    if not args.skip_extras:
        seqs = sharded_sequences(torch, dist, dev, local, rank, world, synth, barrier, max_over_ranks, args.sequences, args.seq_frames)

    out = {
        "metric": METRIC, "value": value, "unit": "features/s", "n_gpus": world, "steps": K, "warmup": Wm,
        "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "u8/int16 fixed-point + f32", "data": "synthetic",
        "config": config_dict(P), "tracked_ok": tracked,
        "e2e": {"value": e2e_val, "unit": "features/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                "ms_per_step": e2e_ms, "steps": Ke, "api": "vo_ft_track_batch (host buffers)",
                "ms_per_step_images_resident": e2e_resident_ms,
                "h2d_only_gbs_per_rank": h2d_gbs_all, "h2d_only_gbs_aggregate": float(sum(h2d_gbs_all)),
                "numa_binding": numa_binding, "ceiling_from_h2d": e2e_ceiling, "frac_of_ceiling": e2e_val / e2e_ceiling, "gpu_numa": numa,
                "note": "ceiling = features / max(device-resident step time, image bytes / measured H2D rate with every rank copying)"},
        "sequences": seqs,
        "gpu_launches": int(launches),
        "roofline": {"bound": "hbm", "kernel": "k_klt3<21>", "achieved": achieved, "peak": peak, "unit": "GB/s",
                     "frac": achieved / peak,
                     # dram__bytes_read.sum + dram__bytes_write.sum of ONE k_klt3 launch (64 pairs x 2000 features) from the
                     # committed `ncu --set full` capture profiles/r2_v2_klt3_full_raw.csv (final round-2 binary): 243.22 MB + 5.24 MB
                     "traffic": 248.5e6 if klt_launches_per_step * 64 == P else None, "traffic_source": "profiles/r2_v2_klt3_full_raw.csv",
                     "peak_source": peak_src,
                     "algorithmic_bytes_per_launch": klt_bytes_per_step / klt_launches_per_step,
                     "launch_ms": klt_ms / klt_launches_per_step, "launches_per_step": klt_launches_per_step,
                     "lk_templates": n_templ, "lk_iterations": n_iter},
        "kernel_ms_per_step": {"pyramids+scharr": pyr_ms, "klt": klt_ms},
        "clocks": clk,
    }

    if rank == 0 and world == 1 and not args.skip_extras:
        out.update(side_measurements(ctx, torch, dev, stream, lefts, rights, pts0_h, synth, capi, args.cfg3_frames))
    if rank == 0 and world == 1:
        cores = os.cpu_count() or 1
        v_all, reps, dt = cpu_reference_rate(lefts[0], rights[0], pts0_h[0], cores, args.cpu_budget)
        v_one, reps1, dt1 = cpu_reference_rate(lefts[0], rights[0], pts0_h[0], 1, min(4.0, args.cpu_budget))
        out["cpu_baseline"] = {
            "value": v_all, "unit": "features/s", "cores": cores, "kind": "reference",
            "sample": f"{reps} x (1 stereo pair, {n} features) in {dt:.1f} s: cv2.calcOpticalFlowPyrLK 4.13 "
                      f"(the library call the reference makes, feature_tracker.cpp:29) + track() post-filter",
            "single_thread_value": v_one}
    elif rank == 0:
        out["cpu_baseline"] = None
    if rank == 0:
        print(json.dumps(out))
    ctx.close()
    if world > 1:
        dist.destroy_process_group()


def gpu_numa_info(local):
    """NUMA node / CPU affinity of this rank's GPU (sysfs), and this process's CPU affinity."""
    info = {}
    try:
        import torch
        bus = torch.cuda.get_device_properties(local).pci_bus_id
        dom = torch.cuda.get_device_properties(local).pci_domain_id
        devn = torch.cuda.get_device_properties(local).pci_device_id
        path = f"/sys/bus/pci/devices/{dom:04x}:{bus:02x}:{devn:02x}.0"
        with open(path + "/numa_node") as f:
            info["gpu_numa_node"] = int(f.read().strip())
        with open(path + "/local_cpulist") as f:
            info["gpu_local_cpulist"] = f.read().strip()
    except Exception as e:
        info["error"] = repr(e)[:80]
    try:
        info["process_cpus"] = len(os.sched_getaffinity(0))
        nodes = [d for d in os.listdir("/sys/devices/system/node") if d.startswith("node")]
        info["host_numa_nodes"] = len(nodes)
    except Exception:
        pass
    return info


def bind_to_gpu_numa(local):
    """Run this rank's threads (and first-touch its pinned buffers) on the CPUs local to its GPU, when the box has
    more than one NUMA node.  Returns what was done."""
    try:
        import torch
        pr = torch.cuda.get_device_properties(local)
        path = f"/sys/bus/pci/devices/{pr.pci_domain_id:04x}:{pr.pci_bus_id:02x}:{pr.pci_device_id:02x}.0/local_cpulist"
        with open(path) as f:
            txt = f.read().strip()
        cpus = set()
        for part in txt.split(","):
            if "-" in part:
                a, b = part.split("-")
                cpus.update(range(int(a), int(b) + 1))
            elif part:
                cpus.add(int(part))
        allowed = os.sched_getaffinity(0)
        cpus &= allowed
        if cpus and cpus != allowed:
            os.sched_setaffinity(0, cpus)
            return f"bound to {len(cpus)} CPUs local to GPU {local}"
        return "GPU-local CPUs == all allowed CPUs (single NUMA node): nothing to bind"
    except Exception as e:
        return "not bound: " + repr(e)[:80]


def sharded_sequences(torch, dist, dev, local, rank, world, synth, barrier, max_over_ranks, n_seq_total, n_frames):
    """BASELINE config 5 for real: n_seq_total independent stereo sequences (full StereoVO::trackStereoImages drop-in: tracking,
    pose GN, detection, new features every frame; reconstruction + local BA on keyframes), rank r runs
    sharding.shard_range(n_seq_total, world, r), one StereoVO instance + host thread + CUDA stream per sequence, no data-path
    collective.  The sequences replay 8 distinct renderings (seeds 5000..5007) round-robin -- 64 distinct 100-frame
    renderings would be 6 GB of host images -- which changes nothing for the GPU (every instance does its own uploads and
    keeps its own state).  frames/s = all frames of all ranks / max-over-ranks wall time between two barriers."""
    import threading
    from visual_odometry_ros_b200 import sharding, stereo_vo as svo
    lo, hi = sharding.shard_range(n_seq_total, world, rank)
    K4, Tlr = synth.kitti_K(), synth.kitti_T_lr()
    n_render = min(8, n_seq_total)
    rend = []
    for i in range(n_render):
        L, R, _ = synth.stereo_sequence(n_frames, W, H, K4, seed=5000 + i, device=str(dev))
        rend.append((torch.from_numpy(L).pin_memory().numpy(), torch.from_numpy(R).pin_memory().numpy()))
    mine = list(range(lo, hi))
    vos = [svo.StereoVO(svo.make_parameters(W, H, K4, K4, Tlr, window_size=WIN, max_level=MAXLVL, n_bins_u=64, n_bins_v=32, device=local))
           for _ in mine]
    for vo, sidx in zip(vos, mine):              # first frame + first step outside the timed region (allocations, module load)
        L, R = rend[sidx % n_render]
        vo.trackStereoImages(L[0], R[0], 0.0)
        vo.trackStereoImages(L[1], R[1], 0.1)
    errors = []
    # host threads: at most one per core this rank can count on (every trackStereoImages ends in a stream synchronize, which
    # spins; 64 threads on 16 cores made this figure swing between 2000 and 3600 frames/s from run to run).  A thread owns
    # several sequences and advances them frame by frame in turn.
    cores = len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 8)
    n_threads = max(1, min(len(vos), max(2, cores // max(1, world))))
    start = threading.Barrier(n_threads + 1, timeout=300)
    kf_count = [0] * len(vos)

    def run(t):
        try:
            own = list(range(t, len(vos), n_threads))
            start.wait()
            for k in range(2, n_frames):
                for j in own:
                    L, R = rend[mine[j] % n_render]
                    vos[j].trackStereoImages(L[k], R[k], 0.1 * k)
                    kf_count[j] += vos[j].frame_info()["keyframe"]
        except Exception as e:
            errors.append(repr(e))
    th = [threading.Thread(target=run, args=(t,), daemon=True) for t in range(n_threads)]
    for t in th:
        t.start()
    barrier()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    try:
        start.wait()
    except threading.BrokenBarrierError:
        errors.append("start barrier broken")
    for t in th:
        t.join()
    dt_ms = max_over_ranks((time.perf_counter() - t0) * 1e3)
    barrier()
    for vo in vos:
        vo.close()
    frames_rank = len(vos) * (n_frames - 2)
    if world > 1:
        t = torch.tensor([float(frames_rank), float(sum(kf_count)), float(len(errors))], dtype=torch.float64, device=dev)
        dist.all_reduce(t)
        frames_all, kf_all, nerr = int(t[0].item()), int(t[1].item()), int(t[2].item())
    else:
        frames_all, kf_all, nerr = frames_rank, sum(kf_count), len(errors)
    res = {"sequences_total": n_seq_total, "sequences_per_rank": len(vos), "frames_each": n_frames - 2, "distinct_renderings": n_render,
           "frames_per_s": frames_all / (dt_ms * 1e-3), "ms_per_frame_amortised": dt_ms / max(1, frames_all) * 1.0,
           "keyframes": kf_all, "errors": nerr, "first_error": errors[:1], "host_threads_per_rank": n_threads,
           "what": "config 5: independent StereoVO sequences sharded over the ranks (one instance + stream each, one host thread per "
                   "available core driving its share of the sequences frame by frame), host u8 images in, pose out, keyframes + local BA "
                   "included; wall clock between barriers, max over ranks"}
    if rank == 0 and world == 1:
        res["cpu"] = cpu_sequence_pool(synth, rend[0], n_frames=8)
    return res


def _cpu_seq_worker(arg):
    L, R = arg
    n_frames = len(L)
    import cv2
    cv2.setNumThreads(1)
    from oracle import stereo_vo as osvo
    from visual_odometry_ros_b200 import synth
    K4, Tlr = synth.kitti_K(), synth.kitti_T_lr()
    ora = osvo.StereoVOOracle(W, H, K4, K4, Tlr, osvo.default_params(window_size=WIN, max_level=MAXLVL, n_bins_u=64, n_bins_v=32))
    ora.track(L[0], R[0])
    t0 = time.perf_counter()
    for k in range(1, n_frames):
        ora.track(L[k], R[k])
    return (n_frames - 1), time.perf_counter() - t0


def cpu_sequence_pool(synth, frames, n_frames=8):
    """The CPU side of config 5: the oracle composition of StereoVO::trackStereoImages (cv2 LK + C restatements), one
    single-threaded process per host core, one sequence each; a bounded sample (the first n_frames frames of a rendering
    the GPU side also runs; every process replays the same frames, which changes nothing for its speed)."""
    import multiprocessing as mp
    L = np.ascontiguousarray(frames[0][:n_frames])
    R = np.ascontiguousarray(frames[1][:n_frames])
    try:
        cores = len(os.sched_getaffinity(0))
    except Exception:
        cores = os.cpu_count() or 1
    try:
        t0 = time.perf_counter()
        with mp.get_context("spawn").Pool(cores) as pool:
            r = pool.map(_cpu_seq_worker, [(L, R) for _ in range(cores)], chunksize=1)
        frames = sum(a for a, _ in r)
        slowest = max(b for _, b in r)
        return {"frames_per_s": frames / slowest, "processes": cores, "frames_each": n_frames - 1, "kind": "port (oracle composition, cv2 LK single-threaded per process)",
                "wall_s_incl_process_start": time.perf_counter() - t0}
    except Exception as e:
        return {"error": repr(e)[:200]}


def side_measurements(ctx, torch, dev, stream, lefts, rights, pts0_h, synth, capi, cfg3_frames=1000):
    """pose-GN solves/s (batched) and single-frame ms (KLT temporal + KLT stereo + pose GN)."""
    from oracle import pose as opose
    res = {}
    # ---- batched pose GN: 4096 problems x 500 points (cfg1 scene, per-problem noise)
    nprob, npts = 4096, 500
    s = synth.pose_scene(seed=1001, n=npts)
    rng = np.random.default_rng(5)
    X = np.tile(s["X"][None], (nprob, 1, 1)).astype(np.float32)
    pl = (s["pts_l1"][None] + rng.normal(0, 0.05, (nprob, npts, 2))).astype(np.float32)
    pr = (s["pts_r1"][None] + rng.normal(0, 0.05, (nprob, npts, 2))).astype(np.float32)
    X_d, pl_d, pr_d = (torch.from_numpy(a).to(dev) for a in (X, pl, pr))
    off_d = torch.arange(0, (nprob + 1) * npts, npts, dtype=torch.int32, device=dev)
    T_d = torch.eye(4, device=dev).repeat(nprob, 1, 1).contiguous()
    mask_d = torch.zeros(nprob * npts, dtype=torch.uint8, device=dev)
    it_d = torch.zeros(nprob, dtype=torch.int32, device=dev)
    ok_d = torch.zeros(nprob, dtype=torch.int32, device=dev)
    K4, Tlr = synth.kitti_K(), synth.kitti_T_lr()
    eye = torch.eye(4, device=dev).repeat(nprob, 1, 1).contiguous()

    def solve():
        T_d.copy_(eye)
        ctx.pose_gn_stereo_batch_d(nprob, off_d.data_ptr(), X_d.data_ptr(), pl_d.data_ptr(), pr_d.data_ptr(), K4, K4,
                                   Tlr, 3.0, T_d.data_ptr(), mask_d.data_ptr(), ok_d.data_ptr(), it_d.data_ptr())
    for _ in range(3):
        solve()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    reps = 10
    a.record(stream)
    for _ in range(reps):
        solve()
    b.record(stream)
    torch.cuda.synchronize()
    ms = a.elapsed_time(b) / reps
    t0 = time.perf_counter()
    creps = 0
    while time.perf_counter() - t0 < 2.0:
        opose.pose_gn_stereo(X[creps % nprob], pl[creps % nprob], pr[creps % nprob], K4, K4, Tlr, 3.0, np.eye(4))
        creps += 1
    cpu_rate = creps / (time.perf_counter() - t0)
    iters_total = float(it_d.double().sum().item())
    peak_hbm, _ = load_peak()
    pose_bytes = 28.0 * npts * iters_total            # 12 + 8 + 8 B per point per iteration
    pose_flops = 260.0 * npts * iters_total           # ~260 FLOP per point per iteration (un-fused FP32)
    fp32_peak_nofma = 148 * 128 * 1.965e9 / 1e12      # TFLOP/s of un-fused FP32 (the kernel is compiled -fmad=false)
    res["pose_gn"] = {"solves_per_s": nprob / (ms * 1e-3), "batch": nprob, "points": npts, "ms_per_batch": ms,
                      "mean_iters": float(it_d.float().mean().item()),
                      "cpu_port_solves_per_s_1core": cpu_rate, "mode": "VO_POSE_FAST (FP64 tree sums)"}
    res["roofline_pose"] = {"kernel": "k_pose_gn<32>", "bound": "hbm", "achieved": pose_bytes / (ms * 1e-3) / 1e9, "peak": peak_hbm, "unit": "GB/s",
                            "frac": pose_bytes / (ms * 1e-3) / 1e9 / peak_hbm,
                            "flops": {"achieved_tflops": pose_flops / (ms * 1e-3) / 1e12, "peak_tflops_fp32_unfused": fp32_peak_nofma,
                                      "frac": pose_flops / (ms * 1e-3) / 1e12 / fp32_peak_nofma},
                            "algorithmic_bytes_per_launch": pose_bytes, "note": "every problem re-reads its 28 B/point each GN iteration from L1 "
                            "(ncu: 94 % L1 hit rate, DRAM 1.4 %, profiles/r2_pose_batch_details.txt); the kernel is bound by instruction issue "
                            "and by the per-iteration serial 6x6 solve, see the flops fraction"}
    # the same batch in strict-order mode (sequential FP32 sums == the reference's arithmetic)
    def solve_strict():
        T_d.copy_(eye)
        ctx.pose_gn_stereo_batch_d(nprob, off_d.data_ptr(), X_d.data_ptr(), pl_d.data_ptr(), pr_d.data_ptr(), K4, K4,
                                   Tlr, 3.0, T_d.data_ptr(), mask_d.data_ptr(), ok_d.data_ptr(), it_d.data_ptr(), flags=capi.VO_POSE_STRICT)
    solve()
    torch.cuda.synchronize()
    T_fast = T_d.clone()
    for _ in range(2):
        solve_strict()
    torch.cuda.synchronize()
    a.record(stream)
    for _ in range(reps):
        solve_strict()
    b.record(stream)
    torch.cuda.synchronize()
    ms_s = a.elapsed_time(b) / reps
    res["pose_gn_strict"] = {"solves_per_s": nprob / (ms_s * 1e-3), "ms_per_batch": ms_s, "mean_iters": float(it_d.float().mean().item()),
                             "max_abs_dT_fast_vs_strict": float((T_fast - T_d).abs().max().item()),
                             "mode": "VO_POSE_STRICT (sequential FP32 sums in point order: the reference's arithmetic, bit for bit)"}
    # ---- single frame: the device-resident stereo tracking step (S1, stereo_vo.cpp:475-670) through the host C ABI:
    #      H2D of the two new images + landmark state, prior, 2x trackWithPrior, trackWithScale, stereo pose GN,
    #      compactions, D2H of pose + survivors; wall clock incl. every copy and the one synchronisation.
    from oracle import step as ostep
    fp = synth.stereo_frame_pair(seed=3003, n=NFEAT)
    ctx.upload_image(0, fp["L0"])
    args = (fp["pts_l0"], fp["pts_r0"], fp["Xw"], fp["tri"], fp["T_wp"], fp["dT_pc_prev"], K4, K4, Tlr, WIN, MAXLVL, THRES_ERR, 3.0)

    def frame(refine):
        return ctx.stereo_track_step(0, 1, 2, fp["L1"], fp["R1"], *args, do_scale_refine=refine, want_counts=False)
    out_frame = {}
    for name, refine in (("with_scale_refine", True), ("track_plus_pose", False)):
        for _ in range(5):
            frame(refine)
        reps = 50
        t0 = time.perf_counter()
        for _ in range(reps):
            g = frame(refine)
        out_frame[name] = (time.perf_counter() - t0) * 1e3 / reps
    import cv2
    cv2.setNumThreads(os.cpu_count() or 1)
    t0 = time.perf_counter()
    creps = 3
    for _ in range(creps):
        o = ostep.stereo_track_step(fp["L0"], fp["L1"], fp["R1"], *args, do_scale_refine=False)
    cpu_ms = (time.perf_counter() - t0) * 1e3 / creps
    res["single_frame"] = {"ms_per_frame": out_frame["track_plus_pose"], "ms_per_frame_with_scale_refine": out_frame["with_scale_refine"],
                           "features": NFEAT, "survivors": int(len(g["index"])),
                           "cpu_ms_per_frame": cpu_ms, "cpu_cores": os.cpu_count() or 1,
                           "what": "vo_stereo_track_step (host buffers): 2 image uploads + prior + 2x trackWithPrior (2000 feat, "
                                   "4 levels, win 21) [+ trackWithScale] + stereo pose GN + compactions, wall clock incl. all "
                                   "copies and the single sync; cpu = the oracle composition (cv2 LK on all threads + C restatements)"}
    res["sequence"] = sequence_measurement(torch, dev, synth, n_frames=cfg3_frames)
    res["sequence_yaml_defaults"] = sequence_measurement(torch, dev, synth, n_frames=min(cfg3_frames, 300), n_cpu=0, with_concurrent=False,
                                                           pose_strict=True)
    res["sequence_all_reference_quirks"] = sequence_measurement(torch, dev, synth, n_frames=min(cfg3_frames, 300), n_cpu=0, with_concurrent=False,
                                                                 pose_strict=True, scale_faithful=True)
    res["mono_sequence"] = mono_sequence_measurement(torch, dev, synth)
    res["sequence_orb"] = sequence_measurement(torch, dev, synth, n_frames=60, n_cpu=6, detector="orb", with_concurrent=False)
    res["lba_depthfilter"] = cfg4_measurement(ctx, synth)
    return res


def concurrent_sequences(svo, Lp, Rp, K4, Tlr, nbu, nbv, n_seq=8, n_frames=60):
    """BASELINE config 5 at sequence level on ONE GPU: n_seq independent StereoVO instances (own context and stream each),
    one host thread per sequence (ctypes releases the GIL inside the C++ step), no inter-sequence traffic.  A single
    sequence is latency-bound (0.6-0.7 ms/frame with the GPU mostly idle); concurrent sequences fill it."""
    import threading
    n_frames = min(n_frames, len(Lp))
    vos = [svo.StereoVO(svo.make_parameters(W, H, K4, K4, Tlr, window_size=WIN, max_level=MAXLVL, n_bins_u=nbu, n_bins_v=nbv)) for _ in range(n_seq)]
    for vo in vos:                       # first frame + first step outside the timed region (allocations)
        vo.trackStereoImages(Lp[0], Rp[0], 0.0)
        vo.trackStereoImages(Lp[1], Rp[1], 0.1)
    bar = threading.Barrier(n_seq + 1, timeout=120)
    errors = []

    def run(vo):
        try:
            bar.wait()
            for k in range(2, n_frames):
                vo.trackStereoImages(Lp[k], Rp[k], 0.1 * k)
            bar.wait()
        except Exception as e:          # a failing sequence must not leave the others (and the bench) waiting
            errors.append(repr(e))
            bar.abort()
    th = [threading.Thread(target=run, args=(vo,), daemon=True) for vo in vos]
    for t in th:
        t.start()
    try:
        bar.wait()
        t0 = time.perf_counter()
        bar.wait()
        dt = time.perf_counter() - t0
    except threading.BrokenBarrierError:
        for t in th:
            t.join(timeout=5)
        return {"sequences": n_seq, "error": errors[:1] or ["barrier broken"]}
    for t in th:
        t.join()
    poses = [vo.pose() for vo in vos]
    same = all(np.array_equal(poses[0], p) for p in poses[1:])
    for vo in vos:
        vo.close()
    frames = n_seq * (n_frames - 2)
    return {"sequences": n_seq, "frames_each": n_frames - 2, "frames_per_s": frames / dt, "ms_per_frame_amortised": dt * 1e3 / frames,
            "identical_results_across_sequences": bool(same)}


def cfg4_measurement(ctx, synth):
    """BASELINE config 4: sliding-window local BA (10 stereo keyframes x 5000 landmarks, Schur-complement LM, 10
    iterations) and the depth filter (20 000 seeds) through the host C ABI, next to the 1-core CPU restatement."""
    from oracle import lba as olba, misc as omisc
    p = synth.lba_problem(seed=4004, n_kf=10, n_points=5000)
    for _ in range(3):
        out = ctx.lba_solve(p)
    reps = 10
    t0 = time.perf_counter()
    for _ in range(reps):
        out = ctx.lba_solve(p)
    gpu_ms = (time.perf_counter() - t0) * 1e3 / reps
    t0 = time.perf_counter()
    rc, poses_o, points_o, avg_o, ok_o = olba.lba_solve(p)
    cpu_ms = (time.perf_counter() - t0) * 1e3
    rng = np.random.default_rng(4004)
    n = 20000
    x0, c0 = rng.uniform(0.02, 0.5, n), rng.uniform(1e-6, 1e-3, n)
    x1, c1 = rng.uniform(0.02, 0.5, n), rng.uniform(1e-6, 1e-3, n)
    for _ in range(3):
        ctx.depth_filter_normal(x0, c0, x1, c1)
    t0 = time.perf_counter()
    for _ in range(20):
        ctx.depth_filter_normal(x0, c0, x1, c1)
    df_ms = (time.perf_counter() - t0) * 1e3 / 20
    t0 = time.perf_counter()
    for _ in range(20):
        omisc.depth_filter_normal(x0, c0, x1, c1)
    df_cpu_ms = (time.perf_counter() - t0) * 1e3 / 20
    # device-resident seeds (vo_depth_filter_normal_d): the update alone, timed with CUDA events on the context's stream
    import torch
    dev = torch.device("cuda", torch.cuda.current_device())
    st = torch.cuda.current_stream()
    xd, cd, md, mc = (torch.from_numpy(a_).to(dev) for a_ in (x0, c0, x1, c1))
    ea, eb = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    for _ in range(3):
        ctx.depth_filter_normal_d(xd.data_ptr(), cd.data_ptr(), md.data_ptr(), mc.data_ptr(), n, xd.data_ptr(), cd.data_ptr())
    ea.record(st)
    for _ in range(100):
        ctx.depth_filter_normal_d(xd.data_ptr(), cd.data_ptr(), md.data_ptr(), mc.data_ptr(), n, xd.data_ptr(), cd.data_ptr())
    eb.record(st)
    torch.cuda.synchronize()
    df_dev_ms = ea.elapsed_time(eb) / 100
    peak_hbm, _ = load_peak()
    n_obs, M, iters = int(p["n_obs"]), int(p["n_points"]), int(p["max_iter"])
    lba_bytes = iters * (185.0 * n_obs + 96.0 * M + 144.0 * n_obs)       # build + per-landmark C,b + Schur re-read of B
    roofline_lba = {"kernel": "k_lba_build + k_lba_reduce + k_lba_solve + k_lba_update_points (x iterations)", "bound": "hbm",
                    "achieved": lba_bytes / (gpu_ms * 1e-3) / 1e9, "peak": peak_hbm, "unit": "GB/s", "frac": lba_bytes / (gpu_ms * 1e-3) / 1e9 / peak_hbm,
                    "algorithmic_bytes_per_call": lba_bytes,
                    "note": "whole vo_lba_solve call incl. H2D/D2H; 10 dependent LM iterations of ~15 MB each: latency-bound (serial 48x48 LDLT per iteration), not HBM-bound"}
    return {"roofline_lba": roofline_lba, "lba_ms": gpu_ms, "lba_cpu_port_ms_1core": cpu_ms, "keyframes": 10, "landmarks": int(p["n_points"]), "observations": int(p["n_obs"]),
            "iterations": int(p["max_iter"]), "lba_final_avg_err_px": float(out[2][-1]), "lba_max_pose_diff_vs_cpu": float(np.abs(out[0] - poses_o).max()),
            "depth_filter_ms": df_ms, "depth_filter_device_resident_ms": df_dev_ms, "depth_filter_cpu_port_ms_1core": df_cpu_ms, "seeds": n,
            "what": "vo_lba_solve / vo_depth_filter_normal with host buffers, wall clock incl. H2D/D2H and the sync"}


def sequence_measurement(torch, dev, synth, n_frames=120, n_cpu=12, detector="harris", with_concurrent=True, pose_strict=False, scale_faithful=False):
    """BASELINE config 3: the full stereo VO step over a synthetic KITTI-like sequence through the reference-API class
    (host images in, pose out; tracking + new features every frame, reconstruction + local BA on keyframes)."""
    from oracle import stereo_vo as osvo
    from visual_odometry_ros_b200 import stereo_vo as svo
    L, R, T_true = synth.stereo_sequence(n_frames, W, H, synth.kitti_K(), seed=3003, device=str(dev))
    K4, Tlr = synth.kitti_K(), synth.kitti_T_lr()
    nbu, nbv = 64, 32
    Lp, Rp = torch.from_numpy(L).pin_memory().numpy(), torch.from_numpy(R).pin_memory().numpy()
    # warm-up instance: first use of every kernel (lazy module loading), first pinned / device allocations
    mk = lambda: svo.StereoVO(svo.make_parameters(W, H, K4, K4, Tlr, window_size=WIN, max_level=MAXLVL, n_bins_u=nbu, n_bins_v=nbv,
                                                  detector=detector, thres_fastscore=20, pose_strict=pose_strict,
                                                  scale_faithful_borders=scale_faithful))
    warm = mk()
    for k in range(min(16, n_frames)):
        warm.trackStereoImages(Lp[k], Rp[k], 0.1 * k)
    warm.close()
    vo = mk()
    ms, kf, nfeat = [], [], []
    parts = {key: [] for key in ("step", "book", "recon", "lba_pack", "lba_solve", "stats", "total")}
    launches0 = vo.launch_count
    for k in range(n_frames):
        t0 = time.perf_counter()
        vo.trackStereoImages(Lp[k], Rp[k], 0.1 * k)
        ms.append((time.perf_counter() - t0) * 1e3)
        fi = vo.frame_info()
        kf.append(fi["keyframe"])
        nfeat.append(fi["n_in"])
        if fi["keyframe"] and k >= 2:
            for key in parts:
                parts[key].append(fi["ms_" + key])
    launches = vo.launch_count - launches0
    T_g = vo.pose()
    gt = np.linalg.inv(T_true[0]) @ T_true[n_frames - 1]
    drift = float(np.linalg.norm(T_g[:3, 3] - gt[:3, 3]) / max(1e-9, np.linalg.norm(gt[:3, 3])))
    vo.close()
    ms, kf = np.asarray(ms[2:]), np.asarray(kf[2:], bool)       # skip the first frame and the first (allocating) step
    import cv2
    cv2.setNumThreads(os.cpu_count() or 1)
    ora = osvo.StereoVOOracle(W, H, K4, K4, Tlr, osvo.default_params(window_size=WIN, max_level=MAXLVL, n_bins_u=nbu, n_bins_v=nbv,
                                                                    detector=detector, fast_threshold=20))
    cms = []
    for k in range(n_cpu):
        t0 = time.perf_counter()
        ora.track(L[k], R[k])
        cms.append((time.perf_counter() - t0) * 1e3)
    conc = concurrent_sequences(svo, Lp, Rp, K4, Tlr, nbu, nbv) if with_concurrent else None
    return {"concurrent_sequences": conc, "detector": "K-det (Harris on the Scharr plane)" if detector == "harris" else "cv::ORB restated (the reference's extractor), FAST threshold 20",
            "pose_mode": "strict (sequential FP32 sums, the reference's arithmetic bit for bit; what the yaml constructor selects)" if pose_strict
                         else "fast (FP64 tree sums; the Parameters-struct default)",
            "scale_border_mode": "reference-faithful stale sample buffers (opt-in: feature_tracker.scale_faithful_borders)" if scale_faithful else "out-of-image samples masked (default)",
            "frames": n_frames, "ms_per_frame_mean": float(ms.mean()), "ms_per_frame_median": float(np.median(ms)),
            "ms_per_non_keyframe": float(ms[~kf].mean()) if (~kf).any() else None,
            "ms_per_keyframe": float(ms[kf].mean()) if kf.any() else None,
            "ms_per_keyframe_median": float(np.median(ms[kf])) if kf.any() else None, "keyframes": int(kf.sum()),
            "keyframe_breakdown_ms": {key: float(np.mean(v)) for key, v in parts.items() if v},
            "mean_tracked_features": float(np.mean(nfeat[2:])), "bins": [nbu, nbv], "gpu_launches": int(launches),
            "translation_drift_vs_ground_truth": drift,
            "cpu_ms_per_frame": float(np.mean(cms[2:])) if len(cms) > 2 else None, "cpu_frames": n_cpu, "cpu_cores": os.cpu_count() or 1,
            "what": "StereoVO::trackStereoImages drop-in (host u8 images in, pose out): fused frame step (pyramids, prior, 2x trackWithPrior, "
                    "trackWithScale, stereo pose GN, compactions, bucketed detection, bidirectional stereo match of new features) every "
                    "frame; reconstruction + 10-iteration local BA on keyframes; wall clock incl. all copies/syncs and the host "
                    "bookkeeping; cpu = oracle composition (cv2 LK on all threads + C restatements + numpy detector)"}


def mono_sequence_measurement(torch, dev, synth, n_frames=120, n_cpu=12):
    """The mono VO step over the left images of the same synthetic sequence through the reference-API class
    MonoVO::trackImage: five-point initialisation on the second image, then tracking + pose-only GN + new features every
    frame, DLT reconstruction + mono local BA on keyframes."""
    from oracle import mono_vo as omvo
    from visual_odometry_ros_b200 import mono_vo as mvo
    L, _, T_true = synth.stereo_sequence(n_frames, W, H, synth.kitti_K(), seed=3003, device=str(dev))
    K4 = synth.kitti_K()
    nbu, nbv = 64, 32
    Lp = torch.from_numpy(L).pin_memory().numpy()
    mk = lambda: mvo.MonoVO(mvo.make_parameters(W, H, K4, window_size=WIN, max_level=MAXLVL, n_bins_u=nbu, n_bins_v=nbv))
    warm = mk()
    for k in range(min(16, n_frames)):
        warm.trackImage(Lp[k], 0.1 * k)
    warm.close()
    vo = mk()
    ms, kf, nfeat, n5 = [], [], [], 0
    launches0 = vo.launch_count
    for k in range(n_frames):
        t0 = time.perf_counter()
        vo.trackImage(Lp[k], 0.1 * k)
        ms.append((time.perf_counter() - t0) * 1e3)
        fi = vo.frame_info()
        kf.append(fi["keyframe"]); nfeat.append(fi["n_in"]); n5 += fi["used_5point"]
    launches = vo.launch_count - launches0
    P = np.stack([vo.frame_pose(j) for j in range(n_frames)])
    T0inv = np.linalg.inv(T_true[0])
    gt = np.stack([T0inv @ T_true[k] for k in range(n_frames)])
    scale = np.linalg.norm(gt[1][:3, 3]) / max(1e-9, np.linalg.norm(P[1][:3, 3]))       # the unit first step fixes the mono scale
    drift = float(np.linalg.norm(scale * P[-1][:3, 3] - gt[-1][:3, 3]) / max(1e-9, np.linalg.norm(gt[-1][:3, 3])))
    ms_init = ms[1]
    vo.close()
    ms, kf = np.asarray(ms[2:]), np.asarray(kf[2:], bool)
    import cv2
    cv2.setNumThreads(os.cpu_count() or 1)
    ora = omvo.MonoVOOracle(W, H, K4, omvo.default_params(window_size=WIN, max_level=MAXLVL, n_bins_u=nbu, n_bins_v=nbv))
    cms = []
    for k in range(n_cpu):
        t0 = time.perf_counter()
        ora.track(L[k])
        cms.append((time.perf_counter() - t0) * 1e3)
    return {"frames": n_frames, "ms_per_frame_mean": float(ms.mean()), "ms_per_frame_median": float(np.median(ms)),
            "ms_per_non_keyframe": float(ms[~kf].mean()) if (~kf).any() else None,
            "ms_per_keyframe": float(ms[kf].mean()) if kf.any() else None, "keyframes": int(kf.sum()),
            "ms_five_point_init_frame": float(ms_init), "five_point_frames": int(n5),
            "mean_tracked_features": float(np.mean(nfeat[2:])), "bins": [nbu, nbv], "gpu_launches": int(launches),
            "scaled_translation_drift_vs_ground_truth": drift,
            "cpu_ms_per_frame": float(np.mean(cms[2:])), "cpu_init_frame_ms": float(cms[1]), "cpu_frames": n_cpu, "cpu_cores": os.cpu_count() or 1,
            "what": "MonoVO::trackImage drop-in (host u8 image in, pose out): five-point RANSAC initialisation (1024 hypotheses) on "
                    "the second image; then the fused mono step (pyramid, prior, trackBidirectionWithPrior, trackWithScale, mono "
                    "pose GN, Sampson gate, compactions, bucketed detection, back-tracking of new features) every frame; DLT "
                    "reconstruction + 10-iteration mono local BA on keyframes; wall clock incl. copies/syncs and host bookkeeping; "
                    "cpu = oracle composition (cv2 LK + cv2.findEssentialMat on all threads + C restatements + numpy detector)"}


if __name__ == "__main__":
    main()
