"""The MonoVO sequence oracle (oracle/mono_vo.py) on the rendered corridor: it initialises with the reference's five-point
call, triangulates, adds keyframes, runs the mono local BA and stays near the rendered truth -- and the host library
exports the MonoVO drop-in with the reference's signatures."""
import ctypes
import os

import numpy as np

from oracle import mono_vo as omvo
from visual_odometry_ros_b200 import synth

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_mono_oracle_sequence():
    W, H = synth.SMALL_W, synth.SMALL_H
    L, _, T = synth.stereo_sequence(10, W, H, synth.small_K(), seed=3103, device="cpu")
    vo = omvo.MonoVOOracle(W, H, synth.small_K(), omvo.default_params(n_bins_u=32, n_bins_v=12, max_level=3, kf_trans=2.0))
    infos = [vo.track(L[k])[1] for k in range(len(L))]
    assert infos[0]["keyframe"] and infos[0]["n_new"] > 200 and not infos[0]["used_5point"]
    assert infos[1]["used_5point"] and infos[1]["n_recon"] > 100                 # mono_vo.cpp:562-696
    assert not any(i["used_5point"] for i in infos[2:])                            # the pose-only BA carries the rest
    assert abs(np.linalg.norm(vo.frames[1].Twc[:3, 3]) - 1.0) < 1e-5               # :606
    assert np.abs(vo.frames[0].dT01[:3, 3] - [0, 0, 1]).max() < 1e-6               # :547-550
    assert sum(i["keyframe"] for i in infos) >= 3
    lbas = [i["lba"] for i in infos if i["lba"] is not None]
    assert lbas and all(l["n_points"] > 50 and l["avg_err"][-1] < 1.0 for l in lbas)
    assert any(i.get("n_recon_kf", 0) > 0 for i in infos)                          # :1032-1076
    # parallax bookkeeping: tracked landmarks of the second image were measured with an identity rotation (:602-611)
    gt = np.linalg.inv(T[0]) @ T[-1]
    P = vo.all_poses()
    s = np.linalg.norm((np.linalg.inv(T[0]) @ T[1])[:3, 3]) / np.linalg.norm(P[1][:3, 3])
    assert np.linalg.norm(s * P[-1][:3, 3] - gt[:3, 3]) < 0.15 * np.linalg.norm(gt[:3, 3])


def test_host_library_exports_mono_vo():
    import subprocess
    host = os.path.join(ROOT, "visual_odometry_ros_b200", "host")
    subprocess.check_call(["make", "-C", host, "-s"])
    hdr = open(os.path.join(host, "mono_vo.h")).read()
    for sig in ("MonoVO(std::string mode, std::string directory_intrinsic);", "void trackImage(const cv::Mat &img, const double &timestamp);",
                "const AlgorithmStatistics &getStatistics() const", "const cv::Mat &getDebugImage()"):
        assert sig in hdr                                                          # mono_vo.h:235-267
    out = subprocess.run(["nm", "-D", "--defined-only", os.path.join(ROOT, "visual_odometry_ros_b200", "libvo_b200_host.so")],
                         capture_output=True, text=True, check=True).stdout
    for sym in ("vo_mvo_create", "vo_mvo_create_from_yaml", "vo_mvo_destroy", "vo_mvo_track", "vo_mvo_pose", "vo_mvo_frame_pose",
                "vo_mvo_frame_info", "vo_mvo_tracks", "vo_mvo_launch_count", "vo_mvo_last_error", "vo_svo_create", "vo_svo_track"):
        assert f" {sym}\n" in out, sym
    # the Python mirror of MonoVO::Parameters / FrameInfo has the C++ layout
    from visual_odometry_ros_b200 import mono_vo as mvo
    from visual_odometry_ros_b200 import stereo_vo as svo
    Hl = ctypes.CDLL(os.path.join(ROOT, "visual_odometry_ros_b200", "libvo_b200_host.so"))
    assert ctypes.sizeof(mvo.Parameters) == Hl.vo_mvo_struct_size(0) and ctypes.sizeof(mvo.FrameInfo) == Hl.vo_mvo_struct_size(1)
    assert ctypes.sizeof(svo.Parameters) == Hl.vo_svo_struct_size(0) and ctypes.sizeof(svo.FrameInfo) == Hl.vo_svo_struct_size(1)
