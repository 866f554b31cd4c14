"""Aggregates an ncu --csv launch list (gpu__time_duration.sum) by kernel and by (kernel, grid)."""
import collections, csv, re, sys
path = sys.argv[1]
top = int(sys.argv[2]) if len(sys.argv) > 2 else 30
with open(path) as f:
    lines = [l for l in f if not l.startswith('==')]
agg = collections.defaultdict(lambda: [0, 0.0])
n = 0
for row in csv.DictReader(lines):
    name = re.sub(r'\(.*', '', row['Kernel Name'])[:70]
    v = float(row['Metric Value'].replace(',', ''))
    u = row['Metric Unit']
    v = v / 1e3 if u == 'ns' else (v * 1e3 if u == 'ms' else v)
    agg[(name, row['Grid Size'])][0] += 1
    agg[(name, row['Grid Size'])][1] += v
    n += 1
tot = sum(v[1] for v in agg.values())
print(f"{n} launches, {tot / 1e3:.2f} ms in total")
byname = collections.defaultdict(lambda: [0, 0.0])
for (nm, g), v in agg.items():
    byname[nm][0] += v[0]; byname[nm][1] += v[1]
for nm, v in sorted(byname.items(), key=lambda kv: -kv[1][1])[:18]:
    print(f"{v[1] / 1e3:9.2f} ms {v[0]:6d}  {nm}")
print()
for (nm, g), v in sorted(agg.items(), key=lambda kv: -kv[1][1])[:top]:
    print(f"{v[1] / 1e3:9.2f} ms {v[0]:5d} avg {v[1] / v[0]:8.1f} us grid {g:>18}  {nm}")
