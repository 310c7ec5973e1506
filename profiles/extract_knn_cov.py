"""Extracts DRAM traffic and warp-instruction counts of the K2+K3 kernel group from an ncu report and writes
profiles/knn_cov_ncu_r2.json together with the SHA-1 of the kernel source it was measured on (bench.py refuses to quote
the numbers for any other version of knn_cov.cu).

    python profiles/extract_knn_cov.py gpurun_out/<report>.ncu-rep      # needs ncu (present in the build container)
"""
import csv
import hashlib
import io
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
rep = sys.argv[1]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr = rows[0]
ix = {h: i for i, h in enumerate(hdr)}
group = ("knn_zero", "brick_clear", "brick_mark", "knn_plan", "knn_lists_tile", "knn_lists_rest", "cov_rest", "cov_from_lists")
kernels = {}
tix, unit_t = ix["gpu__time_duration.sum"], rows[1][ix["gpu__time_duration.sum"]]
for r in rows[2:]:
    name = r[ix["Kernel Name"]].split("(")[0].replace("void ", "")
    base = name.split("<")[0]
    if not any(base.startswith(g) for g in group):
        continue
    t_us = float(r[tix]) * {"ns": 1e-3, "us": 1.0, "ms": 1e3}.get(unit_t, 1.0)
    # the report also holds the scan's (small) launches of the same kernels: the longest launch of each kernel is the
    # 500k-point submap's
    if base in kernels and kernels[base]["time_us"] >= t_us:
        continue
    rd, wr = float(r[ix["dram__bytes_read.sum"]]), float(r[ix["dram__bytes_write.sum"]])
    unit = rows[1][ix["dram__bytes_read.sum"]]
    scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}[unit]
    kernels[base] = {"name": name, "time_us": t_us,
                     "dram_read_MB": rd * scale / 1e6, "dram_write_MB": wr * scale / 1e6,
                     "warp_instructions": float(r[ix["smsp__inst_executed.sum"]]),
                     "issue_active_pct": float(r[ix["smsp__issue_active.avg.pct_of_peak_sustained_active"]]),
                     "warps_active_pct": float(r[ix["sm__warps_active.avg.pct_of_peak_sustained_active"]]),
                     "registers": int(float(r[ix["launch__registers_per_thread"]]))}
kernels = {v.pop("name"): v for v in kernels.values()}
src = os.path.join(ROOT, "direct_lidar_odometry_b200", "csrc", "knn_cov.cu")
out = {"what": "K2+K3 group over the 500 000-point C2 submap (k=20), per launch, ncu --set full --clock-control none",
       "report": os.path.basename(rep), "knn_cov_cu_sha1": hashlib.sha1(open(src, "rb").read()).hexdigest(),
       "kernels": kernels,
       "dram_bytes_per_launch": sum((k["dram_read_MB"] + k["dram_write_MB"]) * 1e6 for k in kernels.values()),
       "warp_instructions_per_launch": sum(k["warp_instructions"] for k in kernels.values()),
       "algorithmic_bytes_per_launch": 64 * 500_000}
json.dump(out, open(os.path.join(ROOT, "profiles", "knn_cov_ncu_r2.json"), "w"), indent=1)
print(json.dumps({k: out[k] for k in ("dram_bytes_per_launch", "warp_instructions_per_launch")}), list(kernels))
