"""Two-GPU landmark-sharded local BA (vo_lba_solve_dist: one ncclAllReduce of the reduced camera system per LM iteration,
sparse_bundle_adjustment.cpp:456-536 partitioned over the landmarks) against the one-GPU solve.  Needs two CUDA devices:
skipped on a one-GPU box (run it with `gpurun --gpus 2`)."""
import json
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_lba_dist_two_gpus_matches_one_gpu():
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
           "--master-port", "29533", os.path.join(ROOT, "tools", "lba_dist_check.py")]
    r = subprocess.run(cmd, cwd=ROOT, capture_output=True, text=True, timeout=600)
    print(r.stdout[-3000:], r.stderr[-2000:])
    assert r.returncode == 0
    lines = [json.loads(l)["lba_dist"] for l in r.stdout.splitlines() if l.startswith('{"lba_dist"')]
    assert len(lines) == 2 and all(l["ok"] for l in lines)
    assert lines[1]["landmarks"] == 100000
