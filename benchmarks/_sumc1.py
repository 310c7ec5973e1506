import json,sys
d=json.load(open(sys.argv[1]))
for k in ("dlo_s2s_cfg","library_defaults"): g=d[k]["gpu"]; print(k, "align_ms", round(g["align_ms"],4), "kernel", round(g["align_kernel_ms"],4), "it", g["iterations"], g["trials"], "cpu", round(d[k]["cpu"]["align_ms"],3), "dpose", d[k]["gpu_vs_cpu_pose"])
