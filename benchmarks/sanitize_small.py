"""Small end-to-end pass over every kernel family for compute-sanitizer (memcheck / racecheck):
voxel filter (exact and speculative pipeline), PointCloud2 decode, index + covariances (warp and tile kNN paths),
fused align with one and two lanes per point, stepped align.
    compute-sanitizer --tool memcheck python benchmarks/sanitize_small.py"""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from direct_lidar_odometry_b200 import NanoGICP, synth, pointcloud2 as pc2  # noqa: E402

T0, T1 = synth.trajectory_pose(0), synth.trajectory_pose(3)
raw0 = synth.crop_box_negative(synth.os1_like(0, T0, beams=32, cols=256))
raw1 = synth.crop_box_negative(synth.os1_like(3, T1, beams=32, cols=256))
g = NanoGICP(0)
v0 = g.voxel_filter(raw0, 0.5)
v1 = g.voxel_filter(raw1, 0.5)                 # second call: speculative single-synchronisation pipeline
n64 = (raw0.shape[0] // 64) * 64
for kind, h, pad in (("ouster", 64, 0), ("velodyne", 1, 0)):
    out = g.preprocess_pointcloud2(pc2.make_pointcloud2(raw0[:n64], kind, height=h, row_pad=pad), 1.0, 0.5)
print("voxel", v0.shape, v1.shape, out.shape)
for k, thr in ((10, 1.0), (20, 0.5)):
    g.setCorrespondenceRandomness(k); g.setMaxCorrespondenceDistance(thr); g.setMaximumIterations(8); g.setTransformationEpsilon(0.01)
    for mode in (0, 1):
        g.setAlignMode(mode)
        g.clearSource(); g.clearTarget()
        g.setInputTarget(v0); g.calculateTargetCovariances()
        g.setInputSource(v1); g.calculateSourceCovariances()
        g.align()
        print("s2s k", k, "mode", mode, "iters", g.result.nr_iterations, g.result.n_compute_error)
# scan-to-map shape: target >= 4x source -> two lanes per source point in the fused kernel
big = np.vstack([v0] + [synth.transform_xyzi(v0, np.array([[1, 0, 0, 0.1 * j], [0, 1, 0, 0.07 * j], [0, 0, 1, 0], [0, 0, 0, 1]], dtype=np.float32)) for j in range(1, 6)])
g.setAlignMode(0)
g.clearSource(); g.clearTarget()
g.setInputTarget(np.ascontiguousarray(big)); g.calculateTargetCovariances()
g.setInputSource(v1[: v1.shape[0] // 2]); g.calculateSourceCovariances()
g.align()
print("s2m-like", big.shape, "iters", g.result.nr_iterations)
