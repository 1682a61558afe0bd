"""Per-kernel summary of an ncu launch list (`--metrics gpu__time_duration.sum --csv`).
usage: summarise_launches.py <launches.csv> <title> > summary.md"""
import collections
import csv
import sys

rows = [r for r in csv.reader(open(sys.argv[1])) if len(r) > 10]
ix = {h: i for i, h in enumerate(rows[0])}
agg = collections.OrderedDict()
for r in rows[1:]:
    if r[ix["Metric Name"]] != "gpu__time_duration.sum":
        continue
    name = r[ix["Kernel Name"]].split("(")[0][:60]
    v = float(r[ix["Metric Value"]].replace(",", ""))
    u = r[ix["Metric Unit"]]
    ms = v / 1e6 if u in ("ns", "nsecond") else v / 1e3 if u.startswith("u") else v
    c = agg.setdefault(name, [0, 0.0])
    c[0] += 1
    c[1] += ms
tot = sum(v[1] for v in agg.values())
print(f"# {sys.argv[2]}\n")
print(f"total device time of all launches: {tot:.1f} ms (cold-cache, serialised; compare shares)\n")
print("| kernel | launches | total ms | share |\n|---|---|---|---|")
for k, (n, ms) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print(f"| `{k}` | {n} | {ms:.2f} | {100 * ms / tot:.1f}% |")
