import sys, time, numpy as np
import os; sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from direct_lidar_odometry_b200 import NanoGICP, synth
g = NanoGICP(0); g.setCorrespondenceRandomness(10)
scans = []
vox = NanoGICP(0)
for i in range(6):
    T = synth.trajectory_pose(i)
    scans.append(vox.voxel_filter(synth.crop_box_negative(synth.os1_like(i, T)), 0.25))
print([s.shape[0] for s in scans])
def run(tag, seq):
    out = []
    for s in seq:
        g.clearSource()
        t0 = time.perf_counter(); g.setInputSource(s); g.sync(); t1 = time.perf_counter()
        out.append((round((t1 - t0) * 1e3, 2), round(g.timings()['set_source_ms'], 2)))
    print(tag, out)
run("same array x8", [scans[0]] * 8)
run("cycle arrays", [scans[i % 6] for i in range(12)])
run("truncated sizes", [scans[0][: 20000 + 137 * i] for i in range(10)])
run("alternate src/tgt", [scans[i % 2] for i in range(8)])
# with swaps like S2S
out = []
g.setInputTarget(scans[0])
for i in range(1, 12):
    t0 = time.perf_counter(); g.setInputSource(scans[i % 6]); g.align(); t1 = time.perf_counter(); g.swapSourceAndTarget()
    out.append((round((t1 - t0) * 1e3, 2), round(g.timings()['set_source_ms'], 2), round(g.timings()['align_ms'], 2)))
print("s2s loop", out)
