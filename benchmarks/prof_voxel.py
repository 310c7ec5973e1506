"""Profiling helper: a few preprocess calls (raw 53k-point scan already on the device) through the fused (voxel_path 0)
or the multi-kernel (1) voxel pipeline — the short command ncu wraps."""
import sys
import os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
from direct_lidar_odometry_b200 import NanoGICP, synth  # noqa: E402
path = int(sys.argv[1]) if len(sys.argv) > 1 else 0
s0 = synth.os1_like(0, synth.trajectory_pose(0))
g = NanoGICP(0)
g.setVoxelPath(path)
dev = torch.from_numpy(s0).cuda()
out = torch.zeros((s0.shape[0], 8), dtype=torch.float32, device="cuda")
for _ in range(4):
    g.preprocess(dev, 1.0, 0.25, out=out)
torch.cuda.synchronize()
print(path, g.timings()["voxel_ms"])
