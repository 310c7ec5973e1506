"""Aggregate `ncu --page source --csv --print-source cuda,sass` output into stall samples per source line.
usage: ncu -i X.ncu-rep --page source --csv --print-source cuda,sass | python profiles/ncu_lines.py [top_n] [kernel_index]"""
import csv, sys
from collections import defaultdict
top = int(sys.argv[1]) if len(sys.argv) > 1 else 30
kidx = int(sys.argv[2]) if len(sys.argv) > 2 else 0
rows = list(csv.reader(sys.stdin))
# sections: ("Kernel Name", ...) then per file: ("File Name", f), header, lines...
kern, first_file, fname, hdr = -1, None, None, None
agg, src = {}, {}
for r in rows:
    if not r: continue
    if r[0] == "File Path":
        if first_file is None: first_file = r[1]
        if r[1] == first_file: kern += 1
        fname = r[1].split("/")[-1]; continue
    if r[0] == "Function Name": continue
    if r[0] == "Line No": hdr = r; continue
    if hdr is None or kern != kidx or not r[0].isdigit(): continue
    d = dict(zip(hdr[4:], r[4:]))
    a = {}
    for k in ("# Samples", "Instructions Executed", "stall_long_sb", "stall_short_sb", "stall_barrier", "stall_wait", "stall_no_inst", "stall_lg", "stall_math", "stall_branch_resolving", "stall_mio", "stall_membar", "stall_sleep"):
        try: a[k] = float(d.get(k, 0) or 0)
        except ValueError: a[k] = 0.0
    agg[(fname, int(r[0]))] = a; src[(fname, int(r[0]))] = r[1]
tot = sum(a["# Samples"] for a in agg.values())
print("total samples", tot)
for ln, a in sorted(agg.items(), key=lambda kv: -kv[1]["# Samples"])[:top]:
    st = {k[6:]: int(v) for k, v in a.items() if k.startswith("stall_") and v > 0.05 * a["# Samples"]}
    print(f"{ln[0]}:{ln[1]:4d} {100*a['# Samples']/max(tot,1):5.1f}% inst={int(a['Instructions Executed']):7d} {st} | {src[ln].strip()[:80]}")
