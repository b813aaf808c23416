"""GPU parity of K-tri / K-df / K-prior / K-compact against the oracle (oracle/misc_oracle.c),
through the C ABI.  Reference lines: triangulate_3d.cpp:5-130, depth_filter.cpp:3-46,
feature_tracker.cpp:208-234, landmark.cpp:194-231."""
import numpy as np
import pytest

from visual_odometry_ros_b200 import synth

pytestmark = pytest.mark.gpu


def _stereo_obs(rng, n, noise=0.0):
    X = np.stack([rng.uniform(-12, 12, n), rng.uniform(-3, 2, n), rng.uniform(4, 50, n)], 1)
    K = synth.kitti_K()
    t10 = np.array([-synth.BASELINE_M, 0, 0], np.float32)
    p0 = np.stack([K[0] * X[:, 0] / X[:, 2] + K[2], K[1] * X[:, 1] / X[:, 2] + K[3]], 1)
    Xr = X + t10
    p1 = np.stack([K[0] * Xr[:, 0] / Xr[:, 2] + K[2], K[1] * Xr[:, 1] / Xr[:, 2] + K[3]], 1)
    p0 += rng.normal(0, noise, p0.shape)
    p1 += rng.normal(0, noise, p1.shape)
    return X, p0.astype(np.float32), p1.astype(np.float32), np.eye(3, dtype=np.float32), t10, K


@pytest.mark.parametrize("n,noise", [(2000, 0.0), (2000, 0.3), (1, 0.1), (0, 0.0)])
def test_triangulate_dlt(gpu_ctx, n, noise):
    from oracle import misc
    rng = np.random.default_rng(4004 + n)
    X, p0, p1, R10, t10, K = _stereo_obs(rng, n, noise)
    X0_g, X1_g = gpu_ctx.triangulate_dlt(p0, p1, R10, t10, K, K)
    X0_o, X1_o = misc.triangulate_dlt(p0, p1, R10, t10, K, K)
    assert X0_g.shape == (n, 3)
    if n == 0:
        return
    # identical algorithm, identical FP32 operation order (no FMA on either side): bit-exact
    assert np.array_equal(X0_g, X0_o), f"max diff {np.abs(X0_g - X0_o).max()}"
    assert np.array_equal(X1_g, X1_o)
    if noise == 0.0:
        rel = np.linalg.norm(X0_g - X, axis=1) / np.linalg.norm(X, axis=1)
        assert rel.max() < 1e-3


def test_triangulate_general_pose(gpu_ctx):
    from oracle import misc
    rng = np.random.default_rng(9)
    n = 500
    X = np.stack([rng.uniform(-5, 5, n), rng.uniform(-2, 2, n), rng.uniform(5, 30, n)], 1)
    R10 = synth.so3_exp([0.01, -0.03, 0.02]).astype(np.float32)
    t10 = np.array([0.3, -0.05, -0.8], np.float32)
    K0 = synth.kitti_K()
    K1 = np.array([700.0, 705.0, 600.0, 180.0], np.float32)
    X1 = X @ R10.T.astype(np.float64) + t10
    p0 = np.stack([K0[0] * X[:, 0] / X[:, 2] + K0[2], K0[1] * X[:, 1] / X[:, 2] + K0[3]], 1).astype(np.float32)
    p1 = np.stack([K1[0] * X1[:, 0] / X1[:, 2] + K1[2], K1[1] * X1[:, 1] / X1[:, 2] + K1[3]], 1).astype(np.float32)
    g0, g1 = gpu_ctx.triangulate_dlt(p0, p1, R10, t10, K0, K1)
    o0, o1 = misc.triangulate_dlt(p0, p1, R10, t10, K0, K1)
    assert np.array_equal(g0, o0) and np.array_equal(g1, o1)
    assert (np.linalg.norm(g0 - X, axis=1) / np.linalg.norm(X, axis=1)).max() < 2e-3


def test_depth_filter_normal_bit_exact(gpu_ctx):
    from oracle import misc
    rng = np.random.default_rng(4004)
    n = 20000
    xp, xc = rng.uniform(0.02, 0.5, n), rng.uniform(0.02, 0.5, n)
    cp, cc = rng.uniform(1e-6, 1e-3, n), rng.uniform(1e-6, 1e-3, n)
    xg, cg = gpu_ctx.depth_filter_normal(xp, cp, xc, cc)
    xo, co = misc.depth_filter_normal(xp, cp, xc, cc)
    assert np.array_equal(xg, xo) and np.array_equal(cg, co)


def test_depth_filter_student_t(gpu_ctx):
    from oracle import misc
    rng = np.random.default_rng(4005)
    n = 20000
    x = rng.uniform(0.02, 0.5, n)
    cov = rng.uniform(1e-4, 1e-2, n)
    a, b = np.full(n, 10.0), np.full(n, 10.0)
    lo, hi = np.full(n, 0.01), np.full(n, 1.0)
    xo, co, ao, bo, loo, hio = x.copy(), cov.copy(), a.copy(), b.copy(), lo.copy(), hi.copy()
    xg, cg, ag, bg, log_, hig = x.copy(), cov.copy(), a.copy(), b.copy(), lo.copy(), hi.copy()
    for it in range(10):   # 10 sequential updates per seed (SURVEY 8d cfg 4)
        meas = x + rng.normal(0, 0.01, n)
        mcov = rng.uniform(1e-4, 1e-3, n)
        xo, co, ao, bo, loo, hio = misc.depth_filter_student_t(xo, co, ao, bo, loo, hio, meas, mcov)
        xg, cg, ag, bg, log_, hig = gpu_ctx.depth_filter_student_t(xg, cg, ag, bg, log_, hig, meas, mcov)
    # FP64; only exp() may differ from glibc by an ulp
    for g, o in ((xg, xo), (cg, co), (ag, ao), (bg, bo)):
        assert np.allclose(g, o, rtol=1e-11, atol=0), np.abs(g / o - 1).max()
    assert np.array_equal(log_, loo) and np.array_equal(hig, hio)


def test_depth_filter_device_resident(gpu_ctx):
    """vo_depth_filter_normal_d / _student_t_d: seeds stay in HBM, 10 in-place updates, same results as the host-buffer entries."""
    import torch
    from oracle import misc
    rng = np.random.default_rng(4006)
    n = 20000
    dev = torch.device("cuda:0")
    x, cov = rng.uniform(0.02, 0.5, n), rng.uniform(1e-4, 1e-2, n)
    xo, co = x.copy(), cov.copy()
    ao, bo, loo, hio = np.full(n, 10.0), np.full(n, 10.0), np.full(n, 0.01), np.full(n, 1.0)
    xs, cs = x.copy(), cov.copy()
    xd, cd = torch.from_numpy(x).to(dev), torch.from_numpy(cov).to(dev)
    xt, ct = torch.from_numpy(x).to(dev), torch.from_numpy(cov).to(dev)
    at, bt = torch.full((n,), 10.0, dtype=torch.float64, device=dev), torch.full((n,), 10.0, dtype=torch.float64, device=dev)
    lot, hit = torch.full((n,), 0.01, dtype=torch.float64, device=dev), torch.full((n,), 1.0, dtype=torch.float64, device=dev)
    torch.cuda.synchronize()
    for it in range(10):
        meas, mcov = x + rng.normal(0, 0.01, n), rng.uniform(1e-4, 1e-3, n)
        md, mc = torch.from_numpy(meas).to(dev), torch.from_numpy(mcov).to(dev)
        torch.cuda.synchronize()
        gpu_ctx.depth_filter_normal_d(xd.data_ptr(), cd.data_ptr(), md.data_ptr(), mc.data_ptr(), n, xd.data_ptr(), cd.data_ptr())   # in place
        gpu_ctx.depth_filter_student_t_d(xt.data_ptr(), ct.data_ptr(), at.data_ptr(), bt.data_ptr(), lot.data_ptr(), hit.data_ptr(),
                                         md.data_ptr(), mc.data_ptr(), n, xt.data_ptr(), ct.data_ptr())
        gpu_ctx.synchronize()
        xs, cs = misc.depth_filter_normal(xs, cs, meas, mcov)
        xo, co, ao, bo, loo, hio = misc.depth_filter_student_t(xo, co, ao, bo, loo, hio, meas, mcov)
    assert np.array_equal(xd.cpu().numpy(), xs) and np.array_equal(cd.cpu().numpy(), cs)
    for g, o in ((xt, xo), (ct, co), (at, ao), (bt, bo)):
        assert np.allclose(g.cpu().numpy(), o, rtol=1e-11, atol=0)
    assert np.array_equal(lot.cpu().numpy(), loo) and np.array_equal(hit.cpu().numpy(), hio)


def test_calc_prior_bit_exact(gpu_ctx):
    from oracle import misc
    rng = np.random.default_rng(12)
    n = 2000
    Xw = np.stack([rng.uniform(-12, 12, n), rng.uniform(-3, 2, n), rng.uniform(4, 50, n)], 1).astype(np.float32)
    Xw[::97] = 0.0
    Tw1 = np.eye(4, dtype=np.float32)
    Tw1[:3, :3] = synth.so3_exp([0.002, -0.012, 0.001])
    Tw1[:3, 3] = [0.02, -0.01, 0.85]
    # make X1 exactly zero for some points: Xw = t  (norm == 0 -> keep pts0)
    Xw[5] = Tw1[:3, 3]
    pts0 = rng.uniform(0, 1000, (n, 2)).astype(np.float32)
    g = gpu_ctx.calc_prior(pts0, Xw, Tw1, synth.kitti_K())
    o = misc.calc_prior(pts0, Xw, Tw1, synth.kitti_K())
    assert np.array_equal(g, o, equal_nan=True)


@pytest.mark.parametrize("n", [0, 1, 31, 1024, 1025, 5000])
def test_compact_indexing_bit_exact(gpu_ctx, n):
    from oracle import misc
    rng = np.random.default_rng(n)
    mask = rng.uniform(size=n) < 0.6
    g = gpu_ctx.compact(mask)
    o = misc.compact(mask)
    assert np.array_equal(g, o)
    assert np.array_equal(g, np.flatnonzero(mask))


def test_triangulate_dlt_grouped_equals_per_group_calls(gpu_ctx):
    """vo_triangulate_dlt_grouped (one relative pose per group of points, one launch) == vo_triangulate_dlt per group, bit for
    bit, == the oracle; out-of-range group indices are refused."""
    from oracle import misc
    from visual_odometry_ros_b200 import capi
    rng = np.random.default_rng(77)
    n, G = 900, 5
    X, p0, p1, _, _, K = _stereo_obs(rng, n, 0.2)
    Rs, ts = [], []
    for g in range(G):
        w = rng.normal(0, 0.02, 3)
        th = np.linalg.norm(w)
        k = w / th
        Kx = np.array([[0, -k[2], k[1]], [k[2], 0, -k[0]], [-k[1], k[0], 0]])
        Rs.append((np.eye(3) + np.sin(th) * Kx + (1 - np.cos(th)) * Kx @ Kx).astype(np.float32))
        ts.append(np.array([-0.5 - 0.1 * g, 0.01 * g, 0.02 * g], np.float32))
    group = rng.integers(0, G, n).astype(np.int32)
    X0, X1 = gpu_ctx.triangulate_dlt_grouped(p0, p1, group, np.stack(Rs), np.stack(ts), K, K)
    for g in range(G):
        sel = group == g
        a0, a1 = gpu_ctx.triangulate_dlt(p0[sel], p1[sel], Rs[g], ts[g], K, K)
        o0, o1 = misc.triangulate_dlt(p0[sel], p1[sel], Rs[g], ts[g], K, K)
        assert np.array_equal(X0[sel], a0) and np.array_equal(X1[sel], a1)
        assert np.array_equal(X0[sel], o0) and np.array_equal(X1[sel], o1)
    bad = group.copy()
    bad[3] = G
    with pytest.raises(capi.VoError):
        gpu_ctx.triangulate_dlt_grouped(p0, p1, bad, np.stack(Rs), np.stack(ts), K, K)
