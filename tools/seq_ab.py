"""A/B helper: the sequence-level measurement of bench.py alone (single sequence + concurrent sequences)."""
import json
import sys
import torch
sys.path.insert(0, ".")
import bench
from visual_odometry_ros_b200 import synth

r = bench.sequence_measurement(torch, torch.device("cuda:0"), synth)
c = r["concurrent_sequences"]
print(json.dumps({"non_kf": r["ms_per_non_keyframe"], "kf": r["ms_per_keyframe"], "mean": r["ms_per_frame_mean"],
                  "conc_fps": c.get("frames_per_s"), "launches": r["gpu_launches"]}))
