"""Kernel shares from an `ncu --metrics gpu__time_duration.sum --csv` launch list."""
import csv, sys, re, collections
rows = [r for r in csv.reader(open(sys.argv[1])) if len(r) > 10 and r[0].isdigit()]
tot = collections.defaultdict(lambda: [0, 0.0])
for r in rows:
    name = re.sub(r"\(.*", "", r[4]).replace("<unnamed>::", "").replace("void ", "")
    tot[name][0] += 1
    tot[name][1] += float(r[-1]) / 1e3
total = sum(v[1] for v in tot.values())
print(f"{len(rows)} launches, {total/1e3:.3f} ms")
for k, v in sorted(tot.items(), key=lambda kv: -kv[1][1]):
    print(f"{k:40s} n={v[0]:4d} total={v[1]/1e3:9.3f} ms avg={v[1]/v[0]:9.1f} us share={100*v[1]/total:5.1f}%")
