"""The stereo sequence measurement of bench.py with the reference's extractor (K-orb)."""
import json
import sys
import torch
sys.path.insert(0, ".")
import bench
from visual_odometry_ros_b200 import synth

r = bench.sequence_measurement(torch, torch.device("cuda:0"), synth, n_frames=60, n_cpu=6, detector="orb", with_concurrent=False)
print(json.dumps({k: v for k, v in r.items() if k != "what"}, indent=1))
