"""Where a voxel-filter / preprocess call spends its time: host-side input (pageable / pinned) or device input, host or
device output, fused (voxel_path 0) or multi-kernel (1) pipeline.  Wall time per call and the device-side phase time."""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
from direct_lidar_odometry_b200 import NanoGICP, synth  # noqa: E402

s0 = synth.os1_like(0, synth.trajectory_pose(0))
for path in (1, 0):
    g = NanoGICP(0)
    g.setVoxelPath(path)
    pin = torch.from_numpy(s0).pin_memory()
    dev = pin.cuda()
    out_d = torch.zeros((s0.shape[0], 8), dtype=torch.float32, device="cuda")
    for name, src in (("pageable", s0), ("pinned", pin), ("device", dev)):
        for oname, out in (("device out", out_d), ("host out", None)):
            for _ in range(5):
                g.preprocess(src, 1.0, 0.25, out=out)
            torch.cuda.synchronize()
            t = time.perf_counter()
            for _ in range(50):
                g.preprocess(src, 1.0, 0.25, out=out)
            torch.cuda.synchronize()
            dt = (time.perf_counter() - t) / 50 * 1e3
            print(f"voxel_path={path} n={s0.shape[0]} {name:9s} -> {oname:10s} preprocess {dt:.4f} ms/call, device-timed {g.timings()['voxel_ms']:.4f}")
