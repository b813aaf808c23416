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
                "void triangulateDLT(const PixelVec &pts0, const PixelVec &pts1, const Rot3 &R10, const Pos3 &t10, CameraConstPtr &cam",
                "void extractORBwithBinning_fast(const cv::Mat &img, PixelVec &pts_extracted, bool flag_nonmax);",
                "void initParams(int n_cols, int n_rows, int n_bins_u, int n_bins_v, int THRES_FAST, int radius);",
                "bool calcPose5PointsAlgorithm(const PixelVec &pts0, const PixelVec &pts1, CameraConstPtr &cam, Rot3 &R10_true, Pos3 &t10_true,"):
        assert sig in hdr


@pytest.mark.gpu
def test_shim_end_to_end_on_gpu():
    subprocess.check_call(["make", "-C", HOST, "-s"])
    out = subprocess.run([os.path.join(HOST, "test_shim")], capture_output=True, text=True, timeout=300)
    assert out.returncode == 0 and "SHIM_OK" in out.stdout, out.stdout[-3000:] + out.stderr[-2000:]
