"""Pins the five-point oracle (oracle/five_point.py): the action-matrix minimal solver on exact data, the cv2 RANSAC +
restated decomposition / cheirality vote (motion_estimator.cpp:21-123, 205-263) on the synthetic two-view case."""
import numpy as np

from oracle import five_point as ofp
from visual_odometry_ros_b200 import synth


def _exact_sets(seed, n_sets):
    rng = np.random.default_rng(seed)
    out = []
    for _ in range(n_sets):
        R = synth.so3_exp(rng.normal(0, 0.05, 3))
        t = rng.normal(0, 1, 3)
        t /= np.linalg.norm(t)
        X = np.stack([rng.uniform(-5, 5, 5), rng.uniform(-3, 3, 5), rng.uniform(4, 30, 5)], 1)
        X1 = X @ R.T + t
        q = np.stack([X[:, 0] / X[:, 2], X[:, 1] / X[:, 2], X1[:, 0] / X1[:, 2], X1[:, 1] / X1[:, 2]], 1)
        tx = np.array([[0, -t[2], t[1]], [t[2], 0, -t[0]], [-t[1], t[0], 0]])
        E = tx @ R
        out.append((q, E / np.linalg.norm(E)))
    return out


def test_minimal_solver_contains_truth_and_satisfies_constraints():
    for q, Et in _exact_sets(11, 20):
        S = ofp.minimal_solutions(q)
        assert 1 <= len(S) <= 10
        assert min(min(np.abs(s - Et).max(), np.abs(s + Et).max()) for s in S) < 1e-8
        for s in S:
            assert np.abs(ofp.cv_error(s, q)).max() < 1e-20
            assert abs(np.linalg.det(s)) < 1e-10
            assert np.abs(2 * s @ s.T @ s - np.trace(s @ s.T) * s).max() < 1e-9


def test_reference_call_sequence_recovers_the_motion():
    sc = synth.two_view_scene()
    ok, R, t, X0, mask, E = ofp.calc_pose_5point(sc["pts0"], sc["pts1"], sc["K4"], 1.0)
    assert ok
    ang = np.arccos(np.clip((np.trace(R.astype(np.float64) @ sc["R10"].T) - 1) / 2, -1, 1))
    # a minimal-sample model without refit (as the reference uses it): a few mrad / a few degrees of direction
    assert ang < 8e-3
    assert np.arccos(np.clip(float(t @ sc["t10"]), -1, 1)) < 6e-2
    inl = np.ones(len(mask), bool)
    inl[sc["outlier_idx"]] = False
    assert mask[inl].mean() > 0.75 and mask[~inl].mean() < 0.1
    assert np.all(X0[mask, 2] > 0)


def test_decomposition_picks_the_cheirality_winner():
    sc = synth.two_view_scene(seed=7, outlier_frac=0.0, noise_px=0.0)
    t = sc["t10"]
    tx = np.array([[0, -t[2], t[1]], [t[2], 0, -t[0]], [-t[1], t[0], 0]])
    for sign in (1.0, -1.0):
        R, tt, X0, m, counts = ofp.decompose_select(sign * (tx @ sc["R10"]), sc["pts0"], sc["pts1"], sc["K4"])
        assert np.abs(R - sc["R10"]).max() < 1e-4 and np.abs(tt - t).max() < 1e-4
        assert max(counts) == len(m) and sorted(counts)[-2] < len(m) // 2
        scale = np.linalg.norm(synth.so3_exp((0.004, -0.02, 0.003)).T @ np.array([0.05, -0.02, 0.9]))   # |t10| of the scene
        assert np.abs(X0 * scale - sc["X0"]).max() < 0.05 * np.abs(sc["X0"]).max()


def test_epipolar_distances_and_one_point_vote():
    """oracle/mono_step.py restatements of motion_estimator.cpp:471-653 on the synthetic two-view case."""
    from oracle import mono_step as omono
    sc = synth.two_view_scene(seed=21, n=2000)
    F = omono.fundamental(sc["K4"], sc["R10"].astype(np.float32), sc["t10"].astype(np.float32))
    ds, de = omono.sampson(sc["pts0"], sc["pts1"], F), omono.symmetric_epipolar(sc["pts0"], sc["pts1"], F)
    inl = np.ones(len(ds), bool)
    inl[sc["outlier_idx"]] = False
    assert np.median(ds[inl]) < 0.5 and np.median(ds[~inl]) > 10 and np.median(de[inl]) < 1.5 and np.median(de[~inl]) > 5
    # (1/g0 + 1/g1)^2 >= 4 / (g0 g1) >= 8 / (g0^2 + g1^2): symmetric^2 >= 8 Sampson, with equality for equal gradients
    assert np.all(de.astype(np.float64) ** 2 >= 8 * ds.astype(np.float64) * (1 - 1e-4))
    assert np.median(de[~inl] ** 2 / (8 * ds[~inl])) < 1.2
    yaw = 0.05
    sp = synth.two_view_scene(seed=33, n=2000, rotvec=(0.0, yaw, 0.0), t=(np.sin(yaw / 2) * 0.9, 0.0, np.cos(yaw / 2) * 0.9), outlier_frac=0.2)
    th, mask, theta, counts, R10, t10 = omono.inliers_1point_histogram(sp["pts0"], sp["pts1"], sp["K4"], 5.0)
    assert abs(th + yaw) <= 1.5 / 400 and counts.sum() <= len(theta) and counts.max() > 0.2 * len(theta)
    inl = np.ones(len(mask), bool)
    inl[sp["outlier_idx"]] = False
    assert mask[inl].mean() > 0.9 and mask[~inl].mean() < 0.3
