"""A/B runs of library variants (csrc/build_variant.sh): C2 phase times with each variants/libnanogicp_<name>.so.
    python benchmarks/ab_knn.py [name ...]      (no names: every variant found, plus the regular build)"""
import glob
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CODE = r'''
import sys, json, numpy as np
sys.path.insert(0, %r)
import torch, bench
from direct_lidar_odometry_b200 import NanoGICP
g = NanoGICP(0)
g.setCorrespondenceRandomness(20); g.setMaxCorrespondenceDistance(0.5); g.setMaximumIterations(32); g.setTransformationEpsilon(0.01)
wl = bench.make_workload(lambda p, leaf: g.voxel_filter(p, leaf))
submap = torch.from_numpy(wl["submap"]).cuda(); scan = torch.from_numpy(wl["scan_0"]).cuda()
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
T = []
for i in range(14):
    flush.zero_(); torch.cuda.synchronize()
    res = bench.gpu_step(g, submap, scan, wl["guesses"][0])
    if i >= 4: T.append(g.timings())
print(json.dumps({k: round(float(np.median([t[k] for t in T])), 4) for k in T[0]}))
''' % ROOT

names = sys.argv[1:]
libs = {"regular": None}
for p in sorted(glob.glob(os.path.join(ROOT, "direct_lidar_odometry_b200", "csrc", "variants", "libnanogicp_*.so"))):
    libs[os.path.basename(p)[len("libnanogicp_"):-3]] = p
for name, path in libs.items():
    if names and name not in names:
        continue
    env = dict(os.environ)
    if path:
        env["NGICP_LIB_PATH"] = path
    out = subprocess.run([sys.executable, "-c", CODE], env=env, capture_output=True, text=True, timeout=600)
    line = out.stdout.strip().splitlines()[-1] if out.stdout.strip() else out.stderr[-400:]
    stats = [l for l in out.stderr.splitlines() if "warp-search" in l][-1:] if "NGICP_KNN_STATS" in os.environ else []
    print(f"{name:12s} {line} {' '.join(stats)}", flush=True)
