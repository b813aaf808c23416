"""The C++ shim (visual_odometry_ros_b200/host) keeps the reference's class API above the C ABI."""
import os
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HOST = os.path.join(ROOT, "visual_odometry_ros_b200", "host")


def test_shim_builds_and_declares_reference_signatures():
    subprocess.check_call(["make", "-C", HOST, "-s"])
    assert os.path.exists(os.path.join(HOST, "test_shim"))
    hdr = open(os.path.join(HOST, "vo_shim.h")).read()
    for sig in ("void track(const cv::Mat &img0, const cv::Mat &img1, const PixelVec &pts0, int window_size, int max_pyr_lvl, float thres_err",
                "bool poseOnlyBundleAdjustment_Stereo(const PointVec &X, const PixelVec &pts_l1, const PixelVec &pts_r1, CameraConstPtr &cam_left",
                "bool solveForFiniteIterations(int MAX_ITER);",
                "void triangulateDLT(const PixelVec &pts0, const PixelVec &pts1, const Rot3 &R10, const Pos3 &t10, CameraConstPtr &cam"):
        assert sig in hdr


@pytest.mark.gpu
def test_shim_end_to_end_on_gpu():
    subprocess.check_call(["make", "-C", HOST, "-s"])
    out = subprocess.run([os.path.join(HOST, "test_shim")], capture_output=True, text=True, timeout=300)
    assert out.returncode == 0 and "SHIM_OK" in out.stdout, out.stdout[-3000:] + out.stderr[-2000:]
