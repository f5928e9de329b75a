"""Per-kernel shares of an `ncu --metrics gpu__time_duration.sum --csv` launch list."""
import collections
import csv
import sys

rows = [r for r in csv.reader(open(sys.argv[1])) if len(r) > 5]
hdr, tot, cnt = None, collections.defaultdict(float), collections.Counter()
for r in rows:
    if r[0] == "ID":
        hdr = r
        continue
    if hdr is None:
        continue
    d = dict(zip(hdr, r))
    name = d["Kernel Name"].split("(")[0].replace("void ", "").replace("pq3d::", "")[:60]
    v = float(d["Metric Value"].replace(",", ""))
    u = d["Metric Unit"]
    v = v / 1000 if u.startswith("n") else (v * 1000 if u.startswith("m") else v)
    tot[name] += v
    cnt[name] += 1
T = sum(tot.values())
print(f"{sum(cnt.values())} launches, {T:.1f} us total (serialised, cold-cache: compare SHARES, not absolutes)\n")
for k, v in sorted(tot.items(), key=lambda x: -x[1]):
    print(f"{v:10.1f} us {100 * v / T:5.1f}%  n={cnt[k]:4d}  avg={v / cnt[k]:8.2f} us  {k}")
