"""CPU: the C-ABI library loads, exports every symbol include/vo_b200.h declares, refuses to run without a
GPU (no CPU fallback), and the host-side logic (level clamp, sharding, 2-rank gloo) works."""
import ctypes
import os
import re
import subprocess
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_symbols():
    hdr = open(os.path.join(ROOT, "include", "vo_b200.h")).read()
    return sorted(set(re.findall(r"^VO_API[^;(]*?\b(vo_[a-z0-9_]+)\s*\(", hdr, flags=re.M)))


def test_library_exports_every_declared_symbol():
    from visual_odometry_ros_b200 import capi
    L = capi.lib()
    names = _declared_symbols()
    assert len(names) >= 30
    missing = [n for n in names if not hasattr(L, n)]
    assert not missing, f"symbols declared in include/vo_b200.h but not exported: {missing}"
    assert L.vo_build_info().decode().startswith("vo_b200") and b"sm_100a" in L.vo_build_info()


def test_header_cites_reference_lines():
    hdr = open(os.path.join(ROOT, "include", "vo_b200.h")).read()
    for cite in ("feature_tracker.cpp:13-206", "motion_estimator.cpp:665-861", "triangulate_3d.cpp:5-130",
                 "depth_filter.cpp:3-13", "sparse_bundle_adjustment.cpp:150-768"):
        assert cite in hdr


def test_no_cpu_fallback_without_device():
    from visual_odometry_ros_b200 import capi
    L = capi.lib()
    h = ctypes.c_void_p()
    rc = L.vo_ctx_create(0, 64, 64, 1, 16, None, ctypes.byref(h))
    if rc == capi.VO_OK:          # running on a GPU box
        L.vo_ctx_destroy(h)
    else:
        assert rc == capi.VO_ERR_NO_DEVICE
        assert b"no CPU fallback" in L.vo_status_string(rc)
        with pytest.raises(capi.VoError):
            capi.Context(0, 64, 64, 1, 16)


def test_product_package_does_not_import_the_oracle():
    pkg = os.path.join(ROOT, "visual_odometry_ros_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".cpp", ".h")):
                txt = open(os.path.join(dirpath, f)).read()
                assert "from oracle" not in txt and "import oracle" not in txt and "oracle/" not in txt, f


def test_effective_max_level_matches_opencv_rule():
    from visual_odometry_ros_b200 import capi
    from oracle import klt as oklt
    L = capi.lib()
    import cv2
    for (w, h, win, ml) in [(1241, 376, 21, 6), (1241, 376, 21, 3), (640, 480, 15, 5), (100, 60, 21, 4), (44, 44, 21, 2)]:
        exp = cv2.buildOpticalFlowPyramid(np.zeros((h, w), np.uint8), (win, win), ml, withDerivatives=False)[0]
        assert L.vo_effective_max_level(w, h, win, ml) == exp == oklt.effective_max_level(w, h, win, ml)


def test_shard_range_partitions():
    from visual_odometry_ros_b200.sharding import shard_range
    for n, world in [(64, 1), (64, 2), (64, 8), (10, 4), (3, 8), (0, 2)]:
        spans = [shard_range(n, world, r) for r in range(world)]
        assert spans[0][0] == 0 and spans[-1][1] == n
        assert all(spans[i][1] == spans[i + 1][0] for i in range(world - 1))
        sizes = [b - a for a, b in spans]
        assert max(sizes) - min(sizes) <= 1
    with pytest.raises(ValueError):
        shard_range(4, 2, 2)


_GLOO_WORKER = r'''
import os, sys
sys.path.insert(0, sys.argv[1])
import torch, torch.distributed as dist
from visual_odometry_ros_b200.sharding import shard_range, max_over_ranks
dist.init_process_group("gloo")
rank, world = dist.get_rank(), dist.get_world_size()
lo, hi = shard_range(64, world, rank)
mine = torch.zeros(64, dtype=torch.int32); mine[lo:hi] = 1
dist.all_reduce(mine)                       # every sequence owned by exactly one rank
assert int(mine.min()) == 1 and int(mine.max()) == 1, mine
t = max_over_ranks(1.0 + rank, dist)        # bench contract: max over ranks
assert t == float(world), t
counts = [torch.zeros(1, dtype=torch.int64) for _ in range(world)]
dist.all_gather(counts, torch.tensor([hi - lo]))
assert sum(int(c) for c in counts) == 64
dist.destroy_process_group()
print("ok", rank)
'''


def test_two_rank_gloo_sharding(tmp_path):
    script = tmp_path / "worker.py"
    script.write_text(_GLOO_WORKER)
    env = dict(os.environ, MASTER_ADDR="127.0.0.1")
    out = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2",
                          "--master-addr", "127.0.0.1", "--master-port", "29531", str(script), ROOT],
                         capture_output=True, text=True, timeout=240, env=env)
    assert out.returncode == 0, out.stdout[-2000:] + out.stderr[-2000:]
    assert out.stdout.count("ok") == 2


def _lba_shard_worker(rank, world, port, q):
    import numpy as np
    import torch.distributed as dist
    from visual_odometry_ros_b200 import sharding, synth
    dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world)
    p = synth.lba_problem(seed=11, n_kf=6, n_points=301)
    mine = sharding.split_lba_problem(p, world, rank)
    lo, hi = mine["landmark_range"]
    # the shard is a self-consistent vo_lba_problem: CSR re-based, all frames kept
    ok = mine["obs_ptr"][0] == 0 and mine["obs_ptr"][-1] == mine["n_obs"] and len(mine["obs_ptr"]) == mine["n_points"] + 1
    ok &= mine["n_frames"] == p["n_frames"] and np.array_equal(mine["poses"], p["poses"])
    ok &= np.array_equal(mine["points"], np.asarray(p["points"]).reshape(-1, 3)[lo:hi])
    o0 = int(p["obs_ptr"][lo])
    ok &= np.array_equal(mine["obs_px"], np.asarray(p["obs_px"]).reshape(-1, 2)[o0:o0 + mine["n_obs"]])
    # the NCCL id would travel like this (bytes object from rank 0)
    box = [bytes(range(128)) if rank == 0 else None]
    dist.broadcast_object_list(box, src=0)
    ok &= box[0] == bytes(range(128))
    sizes = [None] * world
    dist.all_gather_object(sizes, (lo, hi, int(mine["n_obs"])))
    if rank == 0:
        cover = sizes[0][0] == 0 and sizes[-1][1] == p["n_points"] and all(sizes[i][1] == sizes[i + 1][0] for i in range(world - 1))
        q.put(bool(ok) and cover and sum(s[2] for s in sizes) == p["n_obs"])
    else:
        q.put(bool(ok))
    dist.destroy_process_group()


def test_lba_landmark_sharding_two_ranks_gloo():
    """Host logic of the landmark-sharded local BA (world size 2, gloo): shards are disjoint, ordered, cover every landmark
    and observation, each is a valid problem on its own, and the communicator id reaches every rank."""
    import multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29611
    ps = [ctx.Process(target=_lba_shard_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in ps:
        p.start()
    res = [q.get(timeout=120) for _ in ps]
    for p in ps:
        p.join(timeout=60)
    assert all(res)
