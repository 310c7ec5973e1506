import sys, time, numpy as np
sys.path.insert(0, "/root/repo")
import torch
from direct_lidar_odometry_b200 import NanoGICP, synth
T0 = synth.trajectory_pose(0)
s0 = synth.os1_like(0, T0)
for path in (1, 0):
    g = NanoGICP(0); g.setVoxelPath(path)
    pin = torch.from_numpy(s0).pin_memory()
    dev = pin.cuda()
    out = torch.zeros((s0.shape[0], 8), dtype=torch.float32, device="cuda")
    for name, src in (("pageable", s0), ("pinned", pin), ("device", dev)):
        for _ in range(5): g.preprocess(src, 1.0, 0.25, out=out)
        torch.cuda.synchronize(); t=time.perf_counter()
        for _ in range(50): g.preprocess(src, 1.0, 0.25, out=out)
        torch.cuda.synchronize(); dt=(time.perf_counter()-t)/50*1e3
        print(f"voxel_path={path} n={s0.shape[0]} {name:9s} preprocess {dt:.4f} ms/call, device-timed {g.timings()['voxel_ms']:.4f}")
