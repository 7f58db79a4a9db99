#!/usr/bin/env python
"""ncu_lines.py <report.ncu-rep> [launch_index] [topN] -> hottest CUDA source lines (stall samples, executed warp instructions)."""
import csv, io, subprocess, sys
path = sys.argv[1]
which = int(sys.argv[2]) if len(sys.argv) > 2 else 0
top = int(sys.argv[3]) if len(sys.argv) > 3 else 25
out = subprocess.run(["ncu", "-i", path, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
# split per kernel: sections start with a "Kernel Name" row
secs, cur = [], None
for r in csv.reader(io.StringIO(out)):
    if r and r[0] in ("Kernel Name", "Function Name"):
        cur = {"name": r[1], "rows": []}; secs.append(cur); continue
    if r and r[0] == "File Path":
        continue
    if cur is not None: cur["rows"].append(r)
if which < 0:
    for i, sc in enumerate(secs):
        n = 0
        for r in sc["rows"][1:]:
            if len(r) > 8 and r[2] == "-":
                try: n += int(r[6])
                except: pass
        print(i, n, sc["name"][:110])
    sys.exit(0)
sec = secs[which]
print(sec["name"][:150])
rows = sec["rows"]
hdr = rows[0]
ix = {h: i for i, h in enumerate(hdr)}
stall_cols = [i for i, h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h]
agg = []; tot_s = 0; tot_i = 0; stalls = {}
for r in rows[1:]:
    if len(r) > 8 and r[2] == "-":
        try: s = int(r[6]); ie = int(r[7])
        except: continue
        agg.append((s, ie, r[0], r[1].strip()[:120])); tot_s += s; tot_i += ie
    elif len(r) > 8:
        for i in stall_cols:
            try: stalls[hdr[i]] = stalls.get(hdr[i], 0) + int(r[i])
            except: pass
print("samples", tot_s, "warp-inst", tot_i)
T = sum(stalls.values()) or 1
print({k: round(100 * v / T, 1) for k, v in sorted(stalls.items(), key=lambda kv: -kv[1])[:8]})
for a in sorted(agg, reverse=True)[:top]:
    print("%6d %10d  L%-4s %s" % a)
