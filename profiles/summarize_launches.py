"""Per-kernel totals of an `ncu --metrics gpu__time_duration.sum --csv` launch list.
usage: python profiles/summarize_launches.py gpurun_out/launches.csv > profiles/launches_summary.txt"""
import csv, sys, collections, re
rows = [r for r in csv.reader(l for l in open(sys.argv[1]) if l.startswith('"'))]
hdr = rows[0]
ki, vi, ui = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
tot, cnt, mx = collections.Counter(), collections.Counter(), collections.Counter()
for r in rows[1:]:
    if len(r) <= vi: continue
    name = re.sub(r"\(.*", "", r[ki]).replace("ngicp::", "")[:44]
    v = float(r[vi].replace(",", "")) * {"ns": 1e-3, "us": 1.0, "ms": 1e3}.get(r[ui], 1.0)
    tot[name] += v; cnt[name] += 1; mx[name] = max(mx[name], v)
s = sum(tot.values())
for k, v in tot.most_common():
    print(f"{k:44s} n={cnt[k]:4d} total={v:10.1f} us  avg={v / cnt[k]:8.1f}  max={mx[k]:8.1f}  share={100 * v / s:5.1f}%")
