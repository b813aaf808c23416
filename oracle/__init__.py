"""CPU oracle for the visual_odometry_ros hot path -- TEST INFRASTRUCTURE ONLY.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` /
``--impl reference`` legs may import this package.  The product
(``visual_odometry_ros_b200``) never does.

Two oracles live here:

* ``oracle.cv2_klt``  -- the very library call the reference makes
  (``cv::calcOpticalFlowPyrLK``; ``core/visual_odometry/feature_tracker.cpp:29,60,69,108,117,186``)
  through the ``cv2`` 4.13.0 wheel, with the reference's post-filters restated.
* ``oracle.lib``      -- ``libvo_oracle.so``: a plain-C, Eigen-free restatement of the
  reference's own arithmetic (pose-only GN, triangulation, LBA, depth filter,
  trackWithScale) plus a scalar restatement of OpenCV's LK used to pin the spec.
"""
import ctypes
import os
import subprocess

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "libvo_oracle.so")


def build(force: bool = False) -> str:
    """Compile the C oracle in place (gcc only; seconds)."""
    srcs = [os.path.join(_HERE, f) for f in os.listdir(_HERE) if f.endswith(".c")]
    stale = (not os.path.exists(_LIB_PATH)) or any(
        os.path.getmtime(s) > os.path.getmtime(_LIB_PATH) for s in srcs)
    if force or stale:
        subprocess.check_call(["make", "-C", _HERE, "-s", "-B" if force else "-s"])
    return _LIB_PATH


_lib = None


def lib() -> ctypes.CDLL:
    global _lib
    if _lib is None:
        build()
        _lib = ctypes.CDLL(_LIB_PATH)
    return _lib
