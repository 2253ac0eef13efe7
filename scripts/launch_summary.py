"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list per kernel (shares, not absolutes:
ncu times are cold-cache and serialised)."""
import collections
import csv
import sys

rows = [r for r in csv.reader(open(sys.argv[1])) if len(r) > 5 and r[0].isdigit()]
agg = collections.defaultdict(lambda: [0, 0.0])
for r in rows:
    name = r[4].split("(")[0].replace("bsfm::ba::", "").replace("bsfm::match::", "")
    v = float(r[-1].replace(",", "")); unit = r[-2]
    v = v / 1000 if unit == "ns" else (v * 1000 if unit == "ms" else v)
    agg[name][0] += 1; agg[name][1] += v
tot = sum(v[1] for v in agg.values())
print(f"# {sys.argv[1]}: {len(rows)} launches, {tot/1000:.3f} ms of kernel time under ncu (cold-cache, serialised)")
print(f"# {'total_us':>10} {'launches':>8} {'us/launch':>10} {'share':>6}  kernel")
for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print(f"  {v[1]:10.1f} {v[0]:8d} {v[1]/v[0]:10.2f} {100*v[1]/tot:5.1f}%  {k[:90]}")
