"""Warp-stall picture of one kernel from an `ncu --set full --import-source on` report: totals per stall reason and the
most-stalled SASS instructions (`ncu -i X.ncu-rep --page source --csv`)."""
import collections
import csv
import subprocess
import sys

rep = sys.argv[1]
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"], capture_output=True,
                     text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr = next(r for r in rows if "Source" in r)
body = rows[rows.index(hdr) + 1:]
src = hdr.index("Source")
stall_cols = [(i, h) for i, h in enumerate(hdr) if h.startswith("stall_") and "not issued" not in h.lower()]
tot = collections.Counter()
per = []
for r in body:
    if len(r) != len(hdr):
        continue
    d = {}
    for i, h in stall_cols:
        try:
            v = int(float(r[i]))
        except ValueError:
            v = 0
        if v:
            d[h] = v
            tot[h] += v
    if d:
        per.append((sum(d.values()), r[src].strip(), sorted(d.items(), key=lambda kv: -kv[1])[:2]))
n = sum(tot.values())
print(f"{rep}: {n} warp samples\n\nstall reason totals:")
for h, v in tot.most_common(10):
    print(f"{v:8d}  {100 * v / n:4.1f}%  {h}")
print("\ntop instructions:")
for v, s, d in sorted(per, reverse=True)[:15]:
    print(f"{v:7d}  {100 * v / n:4.1f}%  {s[:60]:60s} {[(b, a) for a, b in d]}")
