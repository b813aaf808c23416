"""The mono sequence measurement of bench.py alone."""
import json
import sys
import torch
sys.path.insert(0, ".")
import bench
from visual_odometry_ros_b200 import synth

print(json.dumps(bench.mono_sequence_measurement(torch, torch.device("cuda:0"), synth), indent=1))
