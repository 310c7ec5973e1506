import os, sys, time, numpy as np
sys.path.insert(0, '/root/repo')
from benchmarks import configs as cf
from direct_lidar_odometry_b200 import NanoGICP, synth
import torch
_, T, src = cf._gen_dense(0)
src = np.ascontiguousarray(src)
print("dense scan", src.shape, file=sys.stderr)
g = NanoGICP(0); g.setCorrespondenceRandomness(20)
d = torch.from_numpy(src).cuda()
for rep in range(3):
    g.clearSource(); g.setInputSource(d); g.calculateSourceCovariances(); g.sync()
print(g.timings(), g.grid_info(0))
r = np.linalg.norm(src[:, :3], axis=1)
print("range percentiles", np.percentile(r, [50, 90, 99, 99.9, 100]))
