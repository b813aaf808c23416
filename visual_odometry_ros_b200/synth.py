"""Deterministic synthetic inputs for the BASELINE.json configs.

Pure numpy (+ cv2 for resampling); shared by tests/ and bench.py.  Nothing here is
on the product path -- it only manufactures inputs.

Intrinsics are KITTI-00 (``config/stereo/kitti_00_stereo.yaml:11-48`` of the reference).
"""
import numpy as np

KITTI_W, KITTI_H = 1241, 376
FX = FY = 718.856
CX, CY = 607.1928, 185.2157
BASELINE_M = 0.5371657189


def kitti_K():
    return np.array([FX, FY, CX, CY], np.float32)


def kitti_T_lr():
    T = np.eye(4, dtype=np.float32)
    T[0, 3] = BASELINE_M
    return T


def textured_image(rng, w=KITTI_W, h=KITTI_H, noise_sigma=2.0):
    """Band-limited random texture: 3 octaves of bicubic-upsampled uniform noise."""
    import cv2
    acc = np.zeros((h, w), np.float32)
    for octave, amp in ((8, 1.0), (16, 0.7), (32, 0.5), (64, 0.35)):
        gh, gw = max(2, h // (128 // octave * 2)), max(2, w // (128 // octave * 2))
        g = rng.uniform(-1, 1, (gh, gw)).astype(np.float32)
        acc += amp * cv2.resize(g, (w, h), interpolation=cv2.INTER_CUBIC)
    acc = (acc - acc.min()) / (acc.max() - acc.min())
    img = acc * 255.0 + rng.normal(0, noise_sigma, (h, w)).astype(np.float32)
    return np.clip(np.rint(img), 0, 255).astype(np.uint8)


def warp_translate_field(img, dx, dy):
    """out(x, y) = img(x - dx(x,y), y - dy(x,y)) bilinear; dx/dy scalars or HxW fields."""
    import cv2
    h, w = img.shape
    xs, ys = np.meshgrid(np.arange(w, dtype=np.float32), np.arange(h, dtype=np.float32))
    mapx = (xs - dx).astype(np.float32)
    mapy = (ys - dy).astype(np.float32)
    return cv2.remap(img, mapx, mapy, cv2.INTER_LINEAR, borderMode=cv2.BORDER_REFLECT_101)


def disparity_plane(w=KITTI_W, h=KITTI_H, d_top=5.0, d_bottom=60.0):
    """Ground-plane-like inverse-depth field: disparity grows towards the image bottom."""
    ys = np.linspace(0, 1, h, dtype=np.float32)[:, None]
    xs = np.linspace(-1, 1, w, dtype=np.float32)[None, :]
    return (d_top + (d_bottom - d_top) * ys ** 1.5 + 1.5 * xs).astype(np.float32)


def grid_features(rng, n=2000, w=KITTI_W, h=KITTI_H, border=24, nx=50, ny=40):
    """n features on a jittered nx x ny grid, >= border px from the image edge."""
    gx = np.linspace(border + 4, w - border - 4, nx, dtype=np.float32)
    gy = np.linspace(border + 4, h - border - 4, ny, dtype=np.float32)
    xs, ys = np.meshgrid(gx, gy)
    pts = np.stack([xs.ravel(), ys.ravel()], 1)
    pts = pts + rng.uniform(-3, 3, pts.shape).astype(np.float32)
    pts[:, 0] = np.clip(pts[:, 0], border, w - border)
    pts[:, 1] = np.clip(pts[:, 1], border, h - border)
    idx = rng.permutation(len(pts))[:n]
    idx.sort()
    return np.ascontiguousarray(pts[idx], np.float32)


def klt_stereo_case(seed=2002, n=2000, w=KITTI_W, h=KITTI_H):
    """BASELINE config 2: (left, right, next_left, pts0, disparity field, temporal flow)."""
    rng = np.random.default_rng(seed)
    left = textured_image(rng, w, h)
    disp = disparity_plane(w, h)
    # right image: a scene point at left x appears at x - d in the right image.
    right = warp_translate_field(left, -disp, 0.0)
    right = np.clip(right.astype(np.float32) + rng.normal(0, 1.0, right.shape), 0, 255).astype(np.uint8)
    # next-left: forward-motion-like radial flow around the principal point, <= ~6 px
    xs, ys = np.meshgrid(np.arange(w, dtype=np.float32), np.arange(h, dtype=np.float32))
    fx_ = 0.012 * (xs - CX) * (disp / 30.0) + 1.3
    fy_ = 0.012 * (ys - CY) * (disp / 30.0) - 0.7
    nxt = warp_translate_field(left, fx_, fy_)
    nxt = np.clip(nxt.astype(np.float32) + rng.normal(0, 1.0, nxt.shape), 0, 255).astype(np.uint8)
    pts0 = grid_features(rng, n, w, h)
    return dict(left=left, right=right, next_left=nxt, pts0=pts0, disp=disp, flow=(fx_, fy_))


# ------------------------------------------------------------------ pose GN scene (config 1)
def so3_exp(w):
    w = np.asarray(w, np.float64)
    th = np.linalg.norm(w)
    K = np.array([[0, -w[2], w[1]], [w[2], 0, -w[0]], [-w[1], w[0], 0]])
    if th < 1e-12:
        return np.eye(3) + K
    return np.eye(3) + np.sin(th) / th * K + (1 - np.cos(th)) / th ** 2 * (K @ K)


def pose_scene(seed=1001, n=500, noise_px=0.3, outlier_frac=0.10,
               rotvec=(0.002, -0.012, 0.001), t=(0.02, -0.01, 0.85)):
    """BASELINE config 1: 500-point pinhole scene, stereo observations, 10 % gross outliers.

    Returns X (prev-left frame), pts_l1, pts_r1 (current left/right pixels), T01_true
    (pose of current-left in prev-left: X_prev = T01 * X_cur).
    """
    rng = np.random.default_rng(seed)
    X = np.stack([rng.uniform(-12, 12, n), rng.uniform(-3, 2, n), rng.uniform(4, 50, n)], 1)
    R01 = so3_exp(rotvec)
    t01 = np.asarray(t, np.float64)
    T01 = np.eye(4)
    T01[:3, :3] = R01
    T01[:3, 3] = t01
    R10 = R01.T
    t10 = -R01.T @ t01
    Xl = X @ R10.T + t10
    Xr = Xl.copy()
    Xr[:, 0] -= BASELINE_M  # T_rl = inv(T_lr): x_r = x_l - b
    pl = np.stack([FX * Xl[:, 0] / Xl[:, 2] + CX, FY * Xl[:, 1] / Xl[:, 2] + CY], 1)
    pr = np.stack([FX * Xr[:, 0] / Xr[:, 2] + CX, FY * Xr[:, 1] / Xr[:, 2] + CY], 1)
    pl += rng.normal(0, noise_px, pl.shape)
    pr += rng.normal(0, noise_px, pr.shape)
    n_out = int(round(outlier_frac * n))
    idx = rng.choice(n, n_out, replace=False)
    sign = rng.choice([-1.0, 1.0], (n_out, 2))
    pl[idx] += sign * rng.uniform(5, 30, (n_out, 2))
    return dict(X=X.astype(np.float32), pts_l1=pl.astype(np.float32), pts_r1=pr.astype(np.float32),
                T01_true=T01, outlier_idx=np.sort(idx))


def two_view_scene(seed=6006, n=1500, noise_px=0.3, outlier_frac=0.25, rotvec=(0.004, -0.02, 0.003), t=(0.05, -0.02, 0.9)):
    """Mono initialisation case (mono_vo.cpp:562-696): n correspondences of a pinhole scene between two views of the
    KITTI camera, pixel noise and gross outliers.  Returns pts0, pts1, R10, t10 (unit norm; X1 = R10 X0 + t10), K4."""
    rng = np.random.default_rng(seed)
    X0 = np.stack([rng.uniform(-12, 12, n), rng.uniform(-3, 2, n), rng.uniform(4, 50, n)], 1)
    R01 = so3_exp(rotvec)
    t01 = np.asarray(t, np.float64)
    R10 = R01.T
    t10 = -R01.T @ t01
    X1 = X0 @ R10.T + t10
    p0 = np.stack([FX * X0[:, 0] / X0[:, 2] + CX, FY * X0[:, 1] / X0[:, 2] + CY], 1) + rng.normal(0, noise_px, (n, 2))
    p1 = np.stack([FX * X1[:, 0] / X1[:, 2] + CX, FY * X1[:, 1] / X1[:, 2] + CY], 1) + rng.normal(0, noise_px, (n, 2))
    n_out = int(round(outlier_frac * n))
    idx = rng.choice(n, n_out, replace=False)
    p1[idx] += rng.choice([-1.0, 1.0], (n_out, 2)) * rng.uniform(4, 40, (n_out, 2))
    return dict(pts0=p0.astype(np.float32), pts1=p1.astype(np.float32), R10=R10, t10=t10 / np.linalg.norm(t10),
                K4=np.array([FX, FY, CX, CY], np.float32), outlier_idx=np.sort(idx), X0=X0)


# ------------------------------------------------------------------ local BA problem (config 4)
def lba_problem(seed=4004, n_kf=10, n_points=5000, n_fix=2, noise_px=0.3, point_sigma=0.05,
                rot_sigma_deg=0.5, trans_sigma=0.05, pose_scale=10.0, stereo=True, outlier_frac=0.02):
    """Flat sliding-window BA problem in the layout SparseBAParameters::setPosesAndPoints produces
    (ba_solver/sparse_ba_parameters.h:292-465): reference frame = first keyframe, translations and
    points divided by pose_scale, observations per landmark chronological, left then right.

    Returns a dict of numpy arrays matching include/vo_b200.h::vo_lba_problem (+ ground truth).
    """
    rng = np.random.default_rng(seed)
    # ground-truth keyframe poses T_wk (k -> ref): ~1 m forward per keyframe, gentle yaw
    T_wk = []
    for k in range(n_kf):
        T = np.eye(4)
        T[:3, :3] = so3_exp([0.0, 0.01 * k, 0.0])
        T[:3, 3] = [0.02 * k, -0.005 * k, 1.0 * k]
        T_wk.append(T)
    T_kw = [np.linalg.inv(T) for T in T_wk]
    T_lr = np.eye(4)
    T_lr[0, 3] = BASELINE_M
    T_rl = np.linalg.inv(T_lr)

    k0 = rng.integers(0, n_kf - 1, n_points)
    run = rng.integers(2, n_kf + 1, n_points)
    k1 = np.minimum(k0 + run - 1, n_kf - 1)
    Xw = np.stack([rng.uniform(-10, 10, n_points), rng.uniform(-3, 2, n_points),
                   k1 * 1.0 + rng.uniform(5, 45, n_points)], 1)

    obs_ptr = [0]
    obs_frame, obs_right, obs_px = [], [], []
    for i in range(n_points):
        for k in range(k0[i], k1[i] + 1):
            Xl = T_kw[k][:3, :3] @ Xw[i] + T_kw[k][:3, 3]
            cams = [(0, Xl)]
            if stereo:
                cams.append((1, T_rl[:3, :3] @ Xl + T_rl[:3, 3]))
            for right, Xc in cams:
                px = np.array([FX * Xc[0] / Xc[2] + CX, FY * Xc[1] / Xc[2] + CY]) + rng.normal(0, noise_px, 2)
                if rng.uniform() < outlier_frac:
                    px += rng.uniform(-8, 8, 2)
                obs_frame.append(k)
                obs_right.append(right)
                obs_px.append(px)
        obs_ptr.append(len(obs_frame))

    # initial values: perturbed points and optimisable poses
    Xinit = Xw + rng.normal(0, 1, Xw.shape) * (point_sigma * Xw[:, 2:3])
    poses = np.zeros((n_kf, 4, 4))
    opt_index = np.full(n_kf, -1, np.int32)
    for k in range(n_kf):
        T = T_kw[k].copy()
        if k >= n_fix:
            opt_index[k] = k - n_fix
            dR = so3_exp(rng.normal(0, np.deg2rad(rot_sigma_deg), 3))
            T[:3, :3] = dR @ T[:3, :3]
            T[:3, 3] += rng.normal(0, trans_sigma, 3)
        T[:3, 3] /= pose_scale
        poses[k] = T
    T_lr_s = T_lr.copy()
    T_lr_s[:3, 3] /= pose_scale
    gt_poses = np.stack(T_kw)
    gt_poses[:, :3, 3] /= pose_scale
    return dict(n_frames=n_kf, n_opt=n_kf - n_fix, n_points=n_points, n_obs=len(obs_frame),
                poses=np.ascontiguousarray(poses), opt_index=opt_index,
                points=np.ascontiguousarray(Xinit / pose_scale), obs_ptr=np.asarray(obs_ptr, np.int32),
                obs_frame=np.asarray(obs_frame, np.int32), obs_right=np.asarray(obs_right, np.uint8),
                obs_px=np.ascontiguousarray(np.asarray(obs_px, np.float64)),
                K_l=np.array([FX, FY, CX, CY]), K_r=np.array([FX, FY, CX, CY]),
                T_lr=T_lr_s if stereo else np.eye(4), is_stereo=int(stereo), huber=0.5, lam=1e-5, max_iter=10,
                gt_poses=gt_poses, gt_points=Xw / pose_scale)


# ------------------------------------------------------------------ stereo frame pair (configs 3/5)
def _inverse_warp(img, fwd_dx, fwd_dy, iters=3):
    """Image seen after every source pixel p moved to p + fwd(p): out(q) = img(p) with p + fwd(p) = q
    (fixed-point inversion of the forward flow, bilinear resampling)."""
    import cv2
    h, w = img.shape
    xs, ys = np.meshgrid(np.arange(w, dtype=np.float32), np.arange(h, dtype=np.float32))
    px, py = xs.copy(), ys.copy()
    for _ in range(iters):
        fx = cv2.remap(fwd_dx, px, py, cv2.INTER_LINEAR, borderMode=cv2.BORDER_REPLICATE)
        fy = cv2.remap(fwd_dy, px, py, cv2.INTER_LINEAR, borderMode=cv2.BORDER_REPLICATE)
        px, py = xs - fx, ys - fy
    return cv2.remap(img, px, py, cv2.INTER_LINEAR, borderMode=cv2.BORDER_REFLECT_101), (px, py)


def stereo_frame_pair(seed=3003, n=2000, w=KITTI_W, h=KITTI_H, rotvec=(0.001, -0.008, 0.0005), t=(0.01, -0.005, 0.9),
                      tri_frac=0.9):
    """Two consecutive stereo frames of a static scene (depth from a ground-plane-like disparity field)
    seen by a KITTI-like rig that moves by T01 = (rotvec, t), plus the landmark state StereoVO would
    hold after the first frame. Returns a dict with images L0,R0,L1,R1, pts_l0, pts_r0, Xw, tri,
    T_wp, dT_pc_prev (a slightly wrong constant-velocity guess), T01_true."""
    import cv2
    rng = np.random.default_rng(seed)
    L0 = textured_image(rng, w, h)
    disp0 = disparity_plane(w, h)
    xs, ys = np.meshgrid(np.arange(w, dtype=np.float32), np.arange(h, dtype=np.float32))
    z0 = (FX * BASELINE_M) / disp0
    X0 = np.stack([(xs - CX) / FX * z0, (ys - CY) / FY * z0, z0], -1).astype(np.float64)      # points in left-0 frame
    R01 = so3_exp(rotvec)
    t01 = np.asarray(t, np.float64)
    X1 = (X0 - t01) @ R01                                                                   # = R01^T (X0 - t01)
    u1 = FX * X1[..., 0] / X1[..., 2] + CX
    v1 = FY * X1[..., 1] / X1[..., 2] + CY
    flow_x, flow_y = (u1 - xs).astype(np.float32), (v1 - ys).astype(np.float32)
    noise = lambda im: np.clip(im.astype(np.float32) + rng.normal(0, 1.0, im.shape), 0, 255).astype(np.uint8)
    R0 = noise(_inverse_warp(L0, -disp0, np.zeros_like(disp0))[0])
    L1f, (sx, sy) = _inverse_warp(L0, flow_x, flow_y)
    L1 = noise(L1f)
    disp1_src = (FX * BASELINE_M / X1[..., 2]).astype(np.float32)                             # disparity after the motion, at source pixels
    disp1 = cv2.remap(disp1_src, sx, sy, cv2.INTER_LINEAR, borderMode=cv2.BORDER_REPLICATE)   # at L1 pixels
    R1 = noise(_inverse_warp(L1f, -disp1, np.zeros_like(disp1))[0])

    pts_l0 = grid_features(rng, n, w, h, border=30)
    d0 = cv2.remap(disp0, pts_l0[:, 0].copy(), pts_l0[:, 1].copy(), cv2.INTER_LINEAR).ravel()
    pts_r0 = (pts_l0 - np.stack([d0, np.zeros_like(d0)], 1)).astype(np.float32)
    zf = FX * BASELINE_M / d0
    Xl0 = np.stack([(pts_l0[:, 0] - CX) / FX * zf, (pts_l0[:, 1] - CY) / FY * zf, zf], 1)
    Xl0 = Xl0 * (1.0 + rng.normal(0, 0.003, (n, 1)))                                         # triangulation noise
    T_wp = np.eye(4)
    T_wp[:3, :3] = so3_exp([0.01, 0.3, -0.02])
    T_wp[:3, 3] = [12.0, -0.4, 35.0]
    Xw = Xl0 @ T_wp[:3, :3].T + T_wp[:3, 3]
    tri = (rng.uniform(size=n) < tri_frac).astype(np.uint8)
    T01 = np.eye(4)
    T01[:3, :3] = R01
    T01[:3, 3] = t01
    dT_prev = np.eye(4)
    dT_prev[:3, :3] = so3_exp(np.asarray(rotvec) * 0.8)
    dT_prev[:3, 3] = t01 * 0.9 + [0.01, 0.0, 0.02]
    return dict(L0=L0, R0=R0, L1=L1, R1=R1, pts_l0=pts_l0, pts_r0=pts_r0, Xw=Xw.astype(np.float32), tri=tri,
                T_wp=T_wp.astype(np.float32), dT_pc_prev=dT_prev.astype(np.float32), T01_true=T01)


# ------------------------------------------------------------------ stereo sequence (configs 3/5)
# A static, fully finite scene -- an infinitely long rectangular corridor (ground, ceiling, two walls), every
# surface carrying a periodic 1/f-noise texture with mip levels -- ray-cast per pixel for a KITTI-like rig that
# drives along it (0.8-1.2 m/frame, yaw rate <= 1.5 deg/frame).  Written with torch tensor ops only so that the
# same code renders on the CPU (tests) or on the GPU (bench.py manufactures 1000 frames in seconds); the
# product never sees anything but the resulting u8 images.
def corridor_trajectory(n_frames, seed=3003):
    """Camera-to-world poses T_wc [n,4,4] float64 (x right, y down, z forward along the corridor)."""
    rng = np.random.default_rng(seed)
    speed = rng.uniform(0.8, 1.2, n_frames)
    ph = rng.uniform(0, 2 * np.pi, 3)
    t = np.arange(n_frames)
    yaw = 0.12 * np.sin(2 * np.pi * t / 60.0 + ph[0]) + 0.05 * np.sin(2 * np.pi * t / 23.0 + ph[1])
    pitch = 0.004 * np.sin(2 * np.pi * t / 17.0 + ph[2])
    roll = 0.003 * np.sin(2 * np.pi * t / 29.0 + ph[0])
    T = np.zeros((n_frames, 4, 4))
    pos = np.zeros(3)
    for k in range(n_frames):
        R = so3_exp([0, yaw[k], 0]) @ so3_exp([pitch[k], 0, 0]) @ so3_exp([0, 0, roll[k]])
        if k > 0:
            pos = pos + R[:, 2] * speed[k] * np.array([1.0, 0.0, 1.0])   # stay at constant height
        T[k, :3, :3] = R
        T[k, :3, 3] = pos
        T[k, 3, 3] = 1.0
    return T


def _fractal_texture(seed, size=1024):
    """Periodic 1/f^0.9 noise, contrast-stretched to [0, 255] (float32 numpy, size x size)."""
    rng = np.random.default_rng(seed)
    white = rng.normal(size=(size, size))
    fy = np.fft.fftfreq(size)[:, None]
    fx = np.fft.fftfreq(size)[None, :]
    f = np.sqrt(fx * fx + fy * fy)
    f[0, 0] = 1.0
    spec = np.fft.fft2(white) / f ** 0.9
    spec[0, 0] = 0.0
    tex = np.real(np.fft.ifft2(spec))
    lo, hi = np.percentile(tex, [1, 99])
    return np.clip((tex - lo) / (hi - lo), 0, 1).astype(np.float32) * 255.0


class CorridorRenderer:
    """render(T_wc) -> (left u8 [h,w], right u8 [h,w]) torch tensors on `device`."""

    HALF_WIDTH, GROUND_Y, CEIL_Y, PPM, LEVELS = 7.0, 1.65, -5.0, 128.0, 7

    def __init__(self, w=KITTI_W, h=KITTI_H, K=None, baseline=BASELINE_M, seed=3003, device="cpu", noise_sigma=1.0):
        import torch
        self.torch, self.device, self.w, self.h, self.baseline = torch, torch.device(device), w, h, baseline
        K = kitti_K() if K is None else np.asarray(K, np.float32)
        self.K = [float(v) for v in K]
        self.noise_sigma, self.seed = noise_sigma, seed
        # mip atlas: every level of every plane flattened into one 1-D tensor, so a pixel's two mip levels are
        # fetched with plain gathers (8 taps per pixel whatever its level)
        chunks, self.atlas_off, self.atlas_n = [], [], []
        pos = 0
        for plane in range(4):
            t = torch.from_numpy(_fractal_texture(seed * 7 + plane)).to(self.device)
            offs, ns = [], []
            for _ in range(self.LEVELS):
                offs.append(pos); ns.append(t.shape[0])
                chunks.append(t.reshape(-1))
                pos += t.numel()
                t = torch.nn.functional.avg_pool2d(t[None, None], 2)[0, 0]
            self.atlas_off.append(offs); self.atlas_n.append(ns)
        self.atlas = torch.cat(chunks)
        self.atlas_off_t = torch.tensor(self.atlas_off, dtype=torch.long, device=self.device)      # [4, LEVELS]
        self.atlas_n_t = torch.tensor(self.atlas_n, dtype=torch.long, device=self.device)
        ys, xs = torch.meshgrid(torch.arange(h, dtype=torch.float64, device=self.device),
                                torch.arange(w, dtype=torch.float64, device=self.device), indexing="ij")
        self.dc = torch.stack([(xs - self.K[2]) / self.K[0], (ys - self.K[3]) / self.K[1], torch.ones_like(xs)], -1)
        self.frame_no = 0

    def _sample(self, plane, u, v, lam):
        """Trilinear sample; plane is a per-pixel long tensor."""
        torch = self.torch
        l0 = torch.clamp(torch.floor(lam), 0, self.LEVELS - 2)
        fr = torch.clamp(lam - l0, 0.0, 1.0)
        l0 = l0.long()
        out = torch.zeros_like(u)
        for dl, wgt in ((0, 1.0 - fr), (1, fr)):
            l = l0 + dl
            n = self.atlas_n_t[plane, l]
            off = self.atlas_off_t[plane, l]
            s = torch.pow(0.5, l.to(u.dtype))
            uu, vv = u * s - 0.5, v * s - 0.5
            u0, v0 = torch.floor(uu), torch.floor(vv)
            au, av = uu - u0, vv - v0
            iu0, iv0 = u0.long() % n, v0.long() % n
            iu1, iv1 = (iu0 + 1) % n, (iv0 + 1) % n
            A = self.atlas
            val = (A[off + iv0 * n + iu0] * (1 - au) * (1 - av) + A[off + iv0 * n + iu1] * au * (1 - av) +
                   A[off + iv1 * n + iu0] * (1 - au) * av + A[off + iv1 * n + iu1] * au * av)
            out = out + wgt * val
        return out

    def _render_one(self, R, o, gen):
        torch = self.torch
        d = self.dc @ torch.as_tensor(R, dtype=torch.float64, device=self.device).T          # world ray directions, d_cam.z == 1
        ox, oy, oz = float(o[0]), float(o[1]), float(o[2])
        inf = torch.full_like(d[..., 0], float("inf"))
        tiny = 1e-12
        t_g = torch.where(d[..., 1] > tiny, (self.GROUND_Y - oy) / d[..., 1], inf)
        t_c = torch.where(d[..., 1] < -tiny, (self.CEIL_Y - oy) / d[..., 1], inf)
        t_r = torch.where(d[..., 0] > tiny, (self.HALF_WIDTH - ox) / d[..., 0], inf)
        t_l = torch.where(d[..., 0] < -tiny, (-self.HALF_WIDTH - ox) / d[..., 0], inf)
        ts = torch.stack([t_g, t_c, t_r, t_l], 0)
        t, which = ts.min(0)
        t = torch.clamp(t, max=5000.0)
        P = torch.stack([ox + t * d[..., 0], oy + t * d[..., 1], oz + t * d[..., 2]], -1)
        lam = torch.log2(torch.clamp(t * self.PPM / self.K[0], min=1.0)) + 0.5
        a = torch.where(which < 2, P[..., 0], P[..., 1])
        u = a * self.PPM + 137.0 * which.to(a.dtype)
        v = P[..., 2] * self.PPM
        img = self._sample(which, u, v, lam)
        noise = torch.randn(img.shape, generator=gen, dtype=torch.float64) * self.noise_sigma
        img = img + noise.to(self.device)
        return torch.clamp(torch.round(img), 0, 255).to(torch.uint8), t

    def render(self, T_wc, want_depth=False):
        torch = self.torch
        T_wc = np.asarray(T_wc, np.float64)
        gen = torch.Generator().manual_seed(self.seed * 100003 + self.frame_no)
        self.frame_no += 1
        R, o = T_wc[:3, :3], T_wc[:3, 3]
        left, depth = self._render_one(R, o, gen)
        right, _ = self._render_one(R, o + R[:, 0] * self.baseline, gen)
        return (left, right, depth) if want_depth else (left, right)


def stereo_sequence(n_frames, w=KITTI_W, h=KITTI_H, K=None, seed=3003, device="cpu"):
    """Returns (lefts u8 [n,h,w], rights u8 [n,h,w], T_wc_true [n,4,4]) as numpy arrays."""
    import torch
    traj = corridor_trajectory(n_frames, seed)
    ren = CorridorRenderer(w, h, K, seed=seed, device=device)
    L = torch.empty((n_frames, h, w), dtype=torch.uint8)
    R = torch.empty((n_frames, h, w), dtype=torch.uint8)
    for k in range(n_frames):
        l, r = ren.render(traj[k])
        L[k], R[k] = l.cpu(), r.cpu()
    return L.numpy(), R.numpy(), traj


SMALL_W, SMALL_H = 640, 192


def small_K():
    """Intrinsics for the 640x192 test-size rig (same field of view as KITTI)."""
    s = SMALL_W / KITTI_W
    return np.array([FX * s, FY * s, (SMALL_W - 1) * 0.5, (SMALL_H - 1) * 0.5], np.float32)
