import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


def _has_gpu():
    try:
        import ctypes
        cudart = None
        from visual_odometry_ros_b200 import capi
        L = capi.lib()
        h = ctypes.c_void_p()
        rc = L.vo_ctx_create(0, 64, 64, 0, 0, None, ctypes.byref(h))
        if rc == 0:
            L.vo_ctx_destroy(h)
            return True
        return False
    except Exception:
        return False


@pytest.fixture(scope="session")
def gpu_ctx():
    """A shared 4-slot KITTI-size context. GPU tests fail loudly if the library is missing."""
    from visual_odometry_ros_b200 import capi
    ctx = capi.Context(device=0, max_w=1241, max_h=376, n_slots=6, max_feat=4096)
    yield ctx
    ctx.close()
