"""GPU parity of K-lba-* (local bundle adjustment) against the FP64 oracle restatement of
ba_solver/sparse_bundle_adjustment.cpp:150-768 (oracle/lba_oracle.c), through the C ABI."""
import numpy as np
import pytest

from visual_odometry_ros_b200 import synth

pytestmark = pytest.mark.gpu


def _rot_angle(Ra, Rb):
    dR = Ra @ Rb.T
    return float(np.arcsin(min(1.0, np.linalg.norm(dR - dR.T) / (2.0 * np.sqrt(2.0)))))


@pytest.mark.parametrize("n_points,n_kf,stereo,iters", [(200, 10, True, 10), (5000, 10, True, 10), (300, 6, False, 10),
                                                          (64, 4, True, 3), (1, 3, True, 2)])
def test_lba_matches_oracle(gpu_ctx, n_points, n_kf, stereo, iters):
    from oracle import lba as olba
    p = synth.lba_problem(seed=4004 + n_points, n_kf=n_kf, n_points=n_points, stereo=stereo)
    p["max_iter"] = iters
    rc, poses_o, pts_o, avg_o, ok_o = olba.lba_solve(p)
    assert rc == 0
    poses_g, pts_g, avg_g, ok_g = gpu_ctx.lba_solve(p)
    dt = np.abs(poses_g[:, :3, 3] - poses_o[:, :3, 3]).max() * 10.0        # un-scale (pose_scale = 10) -> metres
    dr = max(_rot_angle(poses_g[k, :3, :3], poses_o[k, :3, :3]) for k in range(n_kf))
    dx = np.abs(pts_g - pts_o).max() * 10.0
    print(f"M={n_points} kf={n_kf} stereo={stereo}: avg_err gpu {avg_g[-1]:.6f} oracle {avg_o[-1]:.6f}  "
          f"dt={dt:.2e} m dr={dr:.2e} rad dX={dx:.2e} m")
    assert ok_g == ok_o
    assert np.allclose(avg_g, avg_o, rtol=1e-9, atol=1e-12)
    assert dt <= 1e-6 and dr <= 1e-6            # BASELINE.json: poses within 1e-6 m / 1e-6 rad
    assert dx <= 1e-6
    # fixed keyframes must be untouched, bit for bit
    fixed = p["opt_index"] < 0
    assert np.array_equal(poses_g[fixed], p["poses"][fixed])


def test_lba_rejects_non_chronological_observations(gpu_ctx):
    from visual_odometry_ros_b200 import capi
    p = synth.lba_problem(seed=1, n_kf=5, n_points=10, stereo=False)
    # reverse the observation order of landmark 0 (needs >= 2 optimisable observations)
    i = int(np.argmax(np.diff(p["obs_ptr"]) >= 4))
    a, b = p["obs_ptr"][i], p["obs_ptr"][i + 1]
    p["obs_frame"][a:b] = p["obs_frame"][a:b][::-1].copy()
    p["obs_px"][a:b] = p["obs_px"][a:b][::-1].copy()
    with pytest.raises(capi.VoError):
        gpu_ctx.lba_solve(p)
