#!/usr/bin/env python
"""launch_summary.py <launches.csv> -> per-kernel count / mean us / share of the listed launches."""
import collections, csv, re, sys
rows = [r for r in csv.reader(open(sys.argv[1])) if len(r) > 10 and r[0].isdigit()]
agg = collections.OrderedDict()
for r in rows:
    name = re.sub(r"\(.*", "", r[4]).replace("void ", "").replace("gsr::", "").replace("<unnamed>::", "").replace("unnamed>::", "").strip()
    m = re.search(r"<([^>]*)>", r[4])
    if m: name += "<" + m.group(1) + ">"
    agg.setdefault((name, r[8]), []).append(float(r[-1]))
tot = sum(sum(v) for v in agg.values())
print(len(rows), "launches, total %.1f us" % (tot / 1e3))
for (k, grid), v in sorted(agg.items(), key=lambda kv: -sum(kv[1])):
    print("%-58s grid %-14s n=%3d mean %8.1f us  share %5.1f%%" % (k[:58], grid, len(v), sum(v) / len(v) / 1e3, 100 * sum(v) / tot))
