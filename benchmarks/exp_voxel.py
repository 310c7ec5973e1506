import os, sys, time, numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from direct_lidar_odometry_b200 import NanoGICP, synth
g = NanoGICP(0)
T = synth.trajectory_pose(3)
s = synth.crop_box_negative(synth.os1_like(3, T))
sp = torch.from_numpy(s).pin_memory()
sd = torch.from_numpy(s).cuda()
outd = torch.empty((s.shape[0], 8), dtype=torch.float32, device="cuda")
outp = torch.empty((s.shape[0], 8), dtype=torch.float32).pin_memory()
def bench(tag, inp, out):
    for _ in range(5): g.voxel_filter(inp, 0.25, out=out)
    ts, ks = [], []
    for _ in range(30):
        t0 = time.perf_counter(); r = g.voxel_filter(inp, 0.25, out=out); t1 = time.perf_counter()
        ts.append((t1 - t0) * 1e3); ks.append(g.timings()["voxel_ms"])
    print(f"{tag:28s} wall {np.median(ts):.3f} ms  gpu-events {np.median(ks):.3f} ms  m={r.shape[0]} of n={s.shape[0]}")
bench("pageable in, numpy out", s, None)
bench("pinned in, pinned out", sp, outp)
bench("device in, device out", sd, outd)
