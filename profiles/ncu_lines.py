"""Aggregate `ncu --page source --csv --print-source cuda,sass` output into stall samples per source line.
usage: ncu -i X.ncu-rep --page source --csv --print-source cuda,sass | python profiles/ncu_lines.py [top_n] [kernel_index]
kernel_index counts the profiled launches in report order (a new launch starts when the function name changes or a
source file repeats within one function); `python profiles/ncu_lines.py 0` lists them."""
import csv, sys
top = int(sys.argv[1]) if len(sys.argv) > 1 else 30
kidx = int(sys.argv[2]) if len(sys.argv) > 2 else 0
rows = list(csv.reader(sys.stdin))
kern, cur_func, seen_files, fname, hdr, pending_file = -1, None, set(), None, None, None
names, agg, src = [], {}, {}
for r in rows:
    if not r: continue
    if r[0] == "File Path":
        pending_file = r[1]; continue
    if r[0] == "Function Name":
        if r[1] != cur_func or pending_file in seen_files:
            kern += 1; cur_func = r[1]; seen_files = set(); names.append(r[1])
        seen_files.add(pending_file)
        fname = pending_file.split("/")[-1]; continue
    if r[0] == "Line No": hdr = r; continue
    if hdr is None or kern != kidx or not r[0].isdigit(): continue
    d = dict(zip(hdr[4:], r[4:]))
    a = {}
    for k in ("# Samples", "Instructions Executed", "stall_long_sb", "stall_short_sb", "stall_barrier", "stall_wait", "stall_no_inst", "stall_lg", "stall_math", "stall_branch_resolving", "stall_mio", "stall_membar", "stall_sleep"):
        try: a[k] = float(d.get(k, 0) or 0)
        except ValueError: a[k] = 0.0
    agg[(fname, int(r[0]))] = a; src[(fname, int(r[0]))] = r[1]
if top == 0:
    for i, n in enumerate(names): print(i, n[:100])
    sys.exit(0)
tot = sum(a["# Samples"] for a in agg.values())
print("kernel", kidx, names[kidx][:100] if kidx < len(names) else "?", "total samples", tot)
for ln, a in sorted(agg.items(), key=lambda kv: -kv[1]["# Samples"])[:top]:
    st = {k[6:]: int(v) for k, v in a.items() if k.startswith("stall_") and v > 0.05 * a["# Samples"]}
    print(f"{ln[0]}:{ln[1]:4d} {100*a['# Samples']/max(tot,1):5.1f}% inst={int(a['Instructions Executed']):7d} {st} | {src[ln].strip()[:80]}")
