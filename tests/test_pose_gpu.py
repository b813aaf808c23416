"""GPU parity of K-pose (pose-only Gauss-Newton) against the FP32 oracle restatement of
motion_estimator.cpp:665-1088 / standalone motion_estimator.cpp:4-411, through the C ABI.

Two accumulation modes (include/vo_b200.h):
  VO_POSE_STRICT  sequential FP32 sums in point order == the reference's arithmetic.  Bar: SAME stop iteration, same
                  iterates, same pose (bit for bit up to the last-ulp of sin/cos in se3Exp_f: asserted <= 1e-7).
  VO_POSE_FAST    FP64 tree sums.  Measured and bounded, not a parity claim: every iterate within 5e-5 of the oracle's for
                  N >= 100 (iterate by iterate, on the common iterations), inlier masks bit-exact, fixed points (no early
                  stop) within the same bound; the sequential FP32 rounding of the reference is what it does not reproduce.
The sweep is the random-problem set of tools/solver_stress.py (N = 11..5000, 0-30 % outliers, noise 0.05-0.5 px)."""
import numpy as np
import pytest

from visual_odometry_ros_b200 import capi, synth

pytestmark = pytest.mark.gpu

TOL_T = 1e-6    # metres  (BASELINE.json north_star)
TOL_R = 1e-6    # radians
STRICT, NOSTOP = capi.VO_POSE_STRICT, capi.VO_POSE_NO_EARLY_STOP


def rot_angle(Ra, Rb):
    dR = Ra.astype(np.float64) @ Rb.astype(np.float64).T
    # small-angle safe: ||dR - dR^T||_F / (2 sqrt 2) == sin(angle)
    s = np.linalg.norm(dR - dR.T) / (2.0 * np.sqrt(2.0))
    return float(np.arcsin(min(1.0, s)))


def pose_delta(Ta, Tb):
    return (float(np.linalg.norm(Ta[:3, 3].astype(np.float64) - Tb[:3, 3])), rot_angle(Ta[:3, :3], Tb[:3, :3]))


def sweep_cases():
    for npts in (11, 37, 100, 500, 2000, 5000):
        for seed in range(6):
            for outl, noise in ((0.1, 0.3), (0.3, 0.5), (0.0, 0.05)):
                yield npts, seed, outl, noise


def scene(npts, seed, outl, noise):
    return synth.pose_scene(seed=100 * seed + npts, n=npts, outlier_frac=outl, noise_px=noise,
                            rotvec=(0.002 * seed, -0.012 + 0.004 * seed, 0.001), t=(0.02, -0.01 * seed, 0.85))


def test_pose_strict_sweep_stereo(gpu_ctx):
    """108 random stereo problems, strict-order mode: stop iteration, every iterate, pose and mask equal the oracle's."""
    from oracle import pose as opose
    K, Tlr = synth.kitti_K(), synth.kitti_T_lr()
    n_cases = n_bitexact = 0
    worst = 0.0
    for npts, seed, outl, noise in sweep_cases():
        s = scene(npts, seed, outl, noise)
        ok_o, T_o, m_o, it_o, tr_o = opose.pose_gn_stereo(s["X"], s["pts_l1"], s["pts_r1"], K, K, Tlr, 3.0, np.eye(4), want_trace=True)
        ok_g, T_g, m_g, it_g, tr_g = gpu_ctx.pose_gn_stereo(s["X"], s["pts_l1"], s["pts_r1"], K, K, Tlr, 3.0, np.eye(4),
                                                            flags=STRICT, want_trace=True)
        assert ok_g == ok_o
        assert it_g == it_o, f"n={npts} seed={seed} outl={outl}: stop iteration {it_g} vs oracle {it_o}"
        assert np.array_equal(m_g, m_o), "inlier masks must be bit-exact"
        d_tr = float(np.abs(tr_g[:, :16] - tr_o[:, :16]).max())
        d_T = float(np.abs(T_g - T_o).max())
        worst = max(worst, d_tr, d_T)
        assert d_tr <= 1e-7 and d_T <= 1e-7, f"n={npts} seed={seed}: iterates differ by {d_tr:.2e}, pose by {d_T:.2e}"
        # the FP32 error the stop test looks at must repeat bit for bit
        assert np.array_equal(tr_g[:, 16], tr_o[:, 16]), f"n={npts} seed={seed}: err_curr trace differs"
        n_bitexact += np.array_equal(T_g, T_o) and np.array_equal(tr_g, tr_o)
        n_cases += 1
    print(f"strict stereo: {n_cases} problems, identical stop iteration in all, {n_bitexact} bit-identical traces, worst |d| {worst:.2e}")
    assert n_cases == 108
    assert n_bitexact >= 0.95 * n_cases


def test_pose_strict_sweep_mono(gpu_ctx):
    """216 mono runs (both error-accounting variants), strict-order mode."""
    from oracle import pose as opose
    K = synth.kitti_K()
    n_cases = n_bitexact = 0
    for npts, seed, outl, noise in sweep_cases():
        s = scene(npts, seed, outl, noise)
        for variant in (0, 1):
            ok_o, R_o, t_o, m_o, it_o, tr_o = opose.pose_gn_mono(s["X"], s["pts_l1"], K, 5, np.eye(3), np.zeros(3), variant, want_trace=True)
            ok_g, R_g, t_g, m_g, it_g, tr_g = gpu_ctx.pose_gn_mono(s["X"], s["pts_l1"], K, 5, np.eye(3), np.zeros(3), variant,
                                                                   flags=STRICT, want_trace=True)
            assert ok_g == ok_o and it_g == it_o, f"n={npts} seed={seed} v{variant}: iterations {it_g} vs {it_o}"
            assert np.array_equal(m_g, m_o)
            d = max(float(np.abs(R_g - R_o).max()), float(np.abs(t_g - t_o).max()), float(np.abs(tr_g[:, :16] - tr_o[:, :16]).max()))
            assert d <= 1e-7, f"n={npts} seed={seed} v{variant}: {d:.2e}"
            assert np.array_equal(tr_g[:, 16], tr_o[:, 16])
            n_bitexact += np.array_equal(R_g, R_o) and np.array_equal(t_g, t_o) and np.array_equal(tr_g, tr_o)
            n_cases += 1
    print(f"strict mono: {n_cases} runs, {n_bitexact} bit-identical traces")
    assert n_cases == 216 and n_bitexact >= 0.95 * n_cases


def test_pose_fast_iterates_and_fixed_point(gpu_ctx):
    """Default (FP64-tree) mode over the same 108 stereo problems: iterate-by-iterate against the oracle's trace on the common
    iterations, masks bit-exact, and the fixed points (30 iterations without the stop test on both sides) within the bar.
    The deviation that remains -- a different STOP iteration -- is counted and bounded, not hidden."""
    from oracle import pose as opose
    K, Tlr = synth.kitti_K(), synth.kitti_T_lr()
    same_it = n_cases = 0
    worst_iter = worst_fix = worst_stop = 0.0
    for npts, seed, outl, noise in sweep_cases():
        s = scene(npts, seed, outl, noise)
        args = (s["X"], s["pts_l1"], s["pts_r1"], K, K, Tlr, 3.0, np.eye(4))
        ok_o, T_o, m_o, it_o, tr_o = opose.pose_gn_stereo(*args, want_trace=True)
        ok_g, T_g, m_g, it_g, tr_g = gpu_ctx.pose_gn_stereo(*args, flags=capi.VO_POSE_FAST, want_trace=True)
        assert ok_g == ok_o and np.array_equal(m_g, m_o), "inlier masks must be bit-exact"
        k = min(it_o, it_g)
        # iterates: T10 after each update.  The FP64-sum mode does NOT meet north_star's 1e-6: the reference's own iterates
        # carry the rounding of 4N sequential FP32 additions (relative error ~ sqrt(4N) * 6e-8 of sums that multiply steps of
        # up to 0.85 m), which moves them by up to 1.7e-5 at N = 2000 (measured on B200) and more for near-minimal problems.
        # That is why VO_POSE_STRICT exists and is what the parity claim rests on; this test bounds and reports the fast
        # mode's deviation
        d_it = max(max(pose_delta(tr_g[i, :16].reshape(4, 4), tr_o[i, :16].reshape(4, 4))) for i in range(k))
        tol = 5e-5 if npts >= 100 else 5e-4
        assert d_it <= tol, f"n={npts} seed={seed} outl={outl}: iterate deviation {d_it:.2e}"
        if npts >= 500:
            worst_iter = max(worst_iter, d_it)
        same_it += it_g == it_o
        worst_stop = max(worst_stop, max(pose_delta(T_g, T_o)))
        # fixed point: both sides run 30 iterations, stop test off
        _, Tf_o, _, _ = opose.pose_gn_stereo(*args, max_iter=30, no_early_stop=True)
        _, Tf_g, _, _ = gpu_ctx.pose_gn_stereo(*args, flags=capi.VO_POSE_FAST | NOSTOP, max_iter=30)
        d_fix = max(pose_delta(Tf_g, Tf_o))
        assert d_fix <= tol, f"n={npts} seed={seed}: fixed points differ by {d_fix:.2e}"
        if npts >= 500:
            worst_fix = max(worst_fix, d_fix)
        n_cases += 1
    print(f"fast stereo: {n_cases} problems; same stop iteration {same_it}; worst iterate deviation (N>=500) {worst_iter:.2e}, "
          f"fixed point {worst_fix:.2e}; worst final-pose deviation when the stop iteration differs {worst_stop:.2e}")
    assert worst_stop <= 1e-4      # the documented consequence of a different stop iteration


@pytest.mark.parametrize("seed,n,thres", [(1001, 500, 3.0), (7, 2000, 3.0), (11, 64, 1.5), (13, 5, 3.0)])
def test_pose_gn_stereo_default_entry(gpu_ctx, seed, n, thres):
    """vo_pose_gn_stereo (the entry the shim calls) in the context's mode == strict after vo_set_pose_mode(STRICT)."""
    from oracle import pose as opose
    s = synth.pose_scene(seed=seed, n=n)
    K, Tlr = synth.kitti_K(), synth.kitti_T_lr()
    ok_o, T_o, m_o, it_o = opose.pose_gn_stereo(s["X"], s["pts_l1"], s["pts_r1"], K, K, Tlr, thres, np.eye(4))
    gpu_ctx.set_pose_mode(STRICT)
    try:
        ok_g, T_g, m_g, it_g = gpu_ctx.pose_gn_stereo(s["X"], s["pts_l1"], s["pts_r1"], K, K, Tlr, thres, np.eye(4))
    finally:
        gpu_ctx.set_pose_mode(capi.VO_POSE_FAST)
    dt, dr = pose_delta(T_g, T_o)
    print(f"n={n}: iters gpu/oracle {it_g}/{it_o}  dt={dt:.2e} m  dr={dr:.2e} rad  inliers {m_g.sum()}/{m_o.sum()}")
    assert ok_g == ok_o and it_g == it_o
    assert dt <= 1e-7 and dr <= 1e-7
    assert np.array_equal(m_g, m_o), "inlier masks must be bit-exact"


@pytest.mark.parametrize("variant", [0, 1])
def test_pose_gn_mono_default_mode(gpu_ctx, variant):
    from oracle import pose as opose
    s = synth.pose_scene(seed=1001, n=500)
    K = synth.kitti_K()
    ok_o, R_o, t_o, m_o, it_o = opose.pose_gn_mono(s["X"], s["pts_l1"], K, 5, np.eye(3), np.zeros(3), variant)
    ok_g, R_g, t_g, m_g, it_g = gpu_ctx.pose_gn_mono(s["X"], s["pts_l1"], K, 5, np.eye(3), np.zeros(3), variant)
    dt = np.linalg.norm(t_g.astype(np.float64) - t_o)
    dr = rot_angle(R_g, R_o)
    print(f"mono v{variant}: iters gpu/oracle {it_g}/{it_o} dt={dt:.2e} dr={dr:.2e}")
    assert ok_g == ok_o and dt <= TOL_T and dr <= TOL_R
    assert np.array_equal(m_g, m_o)


def test_pose_batched_strict_matches_single(gpu_ctx):
    """vo_pose_gn_stereo_batch_ex_d (one CTA per problem) in strict mode == the oracle, problem by problem."""
    import torch
    from oracle import pose as opose
    K, Tlr = synth.kitti_K(), synth.kitti_T_lr()
    sizes = [500, 37, 1200, 64, 5, 2000, 300, 129]
    scenes = [synth.pose_scene(seed=50 + i, n=n) for i, n in enumerate(sizes)]
    off = np.concatenate([[0], np.cumsum(sizes)]).astype(np.int32)
    dev = torch.device("cuda:0")
    X = torch.from_numpy(np.concatenate([s["X"] for s in scenes])).to(dev)
    pl = torch.from_numpy(np.concatenate([s["pts_l1"] for s in scenes])).to(dev)
    pr = torch.from_numpy(np.concatenate([s["pts_r1"] for s in scenes])).to(dev)
    off_d = torch.from_numpy(off).to(dev)
    T = torch.eye(4, dtype=torch.float32, device=dev).repeat(len(sizes), 1, 1).contiguous()
    mask = torch.zeros(int(off[-1]), dtype=torch.uint8, device=dev)
    ok = torch.zeros(len(sizes), dtype=torch.int32, device=dev)
    it = torch.zeros(len(sizes), dtype=torch.int32, device=dev)
    torch.cuda.synchronize()
    gpu_ctx.pose_gn_stereo_batch_d(len(sizes), off_d.data_ptr(), X.data_ptr(), pl.data_ptr(), pr.data_ptr(), K, K, Tlr, 3.0,
                                   T.data_ptr(), mask.data_ptr(), ok.data_ptr(), it.data_ptr(), flags=STRICT)
    gpu_ctx.synchronize()
    T_h, m_h, it_h = T.cpu().numpy(), mask.cpu().numpy().astype(bool), it.cpu().numpy()
    for i, s in enumerate(scenes):
        ok_o, T_o, m_o, it_o = opose.pose_gn_stereo(s["X"], s["pts_l1"], s["pts_r1"], K, K, Tlr, 3.0, np.eye(4))
        assert it_h[i] == it_o and np.abs(T_h[i] - T_o).max() <= 1e-7
        assert np.array_equal(m_h[off[i]:off[i + 1]], m_o)


def test_pose_size_mismatch_raises(gpu_ctx):
    s = synth.pose_scene(n=10)
    with pytest.raises(capi.VoError):
        gpu_ctx.pose_gn_stereo(s["X"], s["pts_l1"][:5], s["pts_r1"], synth.kitti_K(), synth.kitti_K(),
                               synth.kitti_T_lr(), 3.0, np.eye(4))
