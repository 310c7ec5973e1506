"""OdomNode's OWN method bodies (setInputSources, getNextPose, updateKeyframes, getSubmapKeyframes — cut out of the reference
at build time, tests/cpp/odom_extract) timed on the facade: what a DLO build gets per scan by swapping the library and
nothing else (host clouds, host keyframe vectors, Matrix4d covariances copied back and forth as OdomNode does).
    python benchmarks/odom_text_bench.py [n_scans]"""
import json
import os
import struct
import subprocess
import sys
import tempfile

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from direct_lidar_odometry_b200 import NanoGICP, synth  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 300
exe = os.path.join(ROOT, "tests", "cpp", "odom_extract")
vox = NanoGICP(0)
scans, T0 = [], None
for i in range(n):
    T = synth.trajectory_pose(i)
    if i == 0:
        T0 = T.astype(np.float32)
    scans.append(vox.voxel_filter(synth.crop_box_negative(synth.os1_like(i, T)), 0.25))     # preprocessPoints, done by the caller
with tempfile.TemporaryDirectory() as d:
    path = os.path.join(d, "scans.bin")
    with open(path, "wb") as f:
        f.write(struct.pack("i", len(scans)))
        f.write(np.ascontiguousarray(T0.T).tobytes())
        for s in scans:
            f.write(struct.pack("i", s.shape[0]))
            f.write(np.ascontiguousarray(s).tobytes())
    out = subprocess.run([exe, path], capture_output=True, text=True, timeout=1200)
rows = [json.loads(l) for l in out.stdout.splitlines() if l.startswith("{")]
ms = np.array([r["ms"] for r in rows[5:]])
T_last = np.array(rows[-1]["T"], dtype=np.float64).reshape(4, 4).T
truth = synth.trajectory_pose(n - 1)
print(json.dumps({"config": f"OdomNode's own getNextPose / updateKeyframes / getSubmapKeyframes text on the facade, {n} scans (preprocess outside)",
                  "ms_per_scan_mean": float(ms.mean()), "p50": float(np.percentile(ms, 50)), "p99": float(np.percentile(ms, 99)),
                  "keyframes": rows[-1]["keyframes"], "submap_points_last": rows[-1]["submap_points"],
                  "final_translation_error_m": float(np.linalg.norm(T_last[:3, 3] - truth[:3, 3]))}))
