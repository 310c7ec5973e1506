"""Profiling helper: one C1 S2S pair aligned in STEPPED mode (one launch per phase), so that ncu can capture
linearize_kernel / compute_error_kernel — the same device code the fused persistent kernel runs."""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from benchmarks.configs import gen_scans, configure, S2S, S2M  # noqa: E402
from direct_lidar_odometry_b200 import NanoGICP  # noqa: E402

scans = gen_scans([0, 1])
g = NanoGICP(0)
configure(g, S2S)
g.setAlignMode(1)
v0 = g.voxel_filter(scans[0][1], 0.25); v1 = g.voxel_filter(scans[1][1], 0.25)
for _ in range(3):
    g.clearSource(); g.clearTarget()
    g.setInputTarget(v0); g.calculateTargetCovariances()
    g.setInputSource(v1); g.calculateSourceCovariances()
    g.align()
print(g.result.nr_iterations, g.timings())
