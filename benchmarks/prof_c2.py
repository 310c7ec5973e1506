"""Profiling helper: N plain C2 steps (bench.py's gpu_step on device-resident inputs, no CPU leg, no timing harness) —
the short command ncu wraps.  NGICP_KNN_STATS=1 prints the tile path's item / fallback statistics per step."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
import bench  # noqa: E402
from direct_lidar_odometry_b200 import NanoGICP  # noqa: E402

steps = int(sys.argv[1]) if len(sys.argv) > 1 else 4
g = NanoGICP(0)
g.setCorrespondenceRandomness(bench.S2M["k"]); g.setMaxCorrespondenceDistance(bench.S2M["thr"])
g.setMaximumIterations(bench.S2M["max_iter"]); g.setTransformationEpsilon(bench.S2M["trans_eps"])
if os.environ.get("NGICP_CELL"):
    g.setGridCellSize(float(os.environ["NGICP_CELL"]))   # experiments: fixed cell edge instead of the density-driven one
if os.environ.get("NGICP_INDEX_PATH"):
    g.setIndexPath(int(os.environ["NGICP_INDEX_PATH"]))   # experiments: 0 cooperative, 1 multi-kernel, 3 cluster (small clouds)
wl = bench.make_workload(lambda p, leaf: g.voxel_filter(p, leaf))
submap = torch.from_numpy(wl["submap"]).cuda()
scan = torch.from_numpy(wl["scan_0"]).cuda()
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
T = []
for _ in range(steps):
    flush.zero_(); torch.cuda.synchronize()
    res = bench.gpu_step(g, submap, scan, wl["guesses"][0])
    T.append(g.timings())
if steps >= 8:
    print("median of last", steps - 4, {k: round(float(np.median([t[k] for t in T[4:]])), 4) for k in T[0]})
print(res.nr_iterations, {k: round(v, 4) for k, v in g.timings().items()}, g.grid_info(1))
