"""Latency of setInputTarget (snapshot + search index) per index_path: 0 cooperative fused kernel, 1 multi-kernel pipeline,
3 one thread-block cluster (small clouds).  Device-resident input, wall time per call including the synchronisation."""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
from direct_lidar_odometry_b200 import NanoGICP, synth  # noqa: E402

s0 = synth.os1_like(0, synth.trajectory_pose(0))
vox = NanoGICP(0)
clouds = {"22k voxelised scan": vox.voxel_filter(s0, 0.25), "53k raw scan": s0}
for name, c in clouds.items():
    d = torch.from_numpy(np.ascontiguousarray(c)).cuda()
    for path in (0, 1, 3):
        g = NanoGICP(0)
        g.setIndexPath(path)
        for _ in range(10):
            g.clearTarget(); g.setInputTarget(d); g.sync()
        t = time.perf_counter()
        for _ in range(100):
            g.clearTarget(); g.setInputTarget(d); g.sync()
        dt = (time.perf_counter() - t) / 100 * 1e3
        print(f"{name:20s} n={c.shape[0]:6d} index_path={path}: {dt:.4f} ms/call (device-timed {g.timings()['set_target_ms']:.4f})")
