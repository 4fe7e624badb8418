"""Summarise an ncu `--metrics gpu__time_duration.sum --csv` launch list by kernel and grid size."""
import csv, collections, re, sys
lines = [l for l in open(sys.argv[1]) if not l.startswith('==')]
agg = collections.OrderedDict(); tot = 0
for row in csv.DictReader(lines):
    name = re.sub(r'\(.*', '', row['Kernel Name'])[:64]
    try: t = float(row['Metric Value'].replace(',', ''))
    except Exception: continue
    u = row['Metric Unit']
    t *= {'ns': 1, 'us': 1e3, 'ms': 1e6, 's': 1e9}.get(u, 1)
    a = agg.setdefault((name, row.get('Grid Size', '')), [0, 0.0]); a[0] += 1; a[1] += t; tot += t
top = int(sys.argv[2]) if len(sys.argv) > 2 else 20
for k, (c, t) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:top]:
    print("%-66s grid %-14s n=%4d total %9.3f ms avg %8.1f us %5.1f%%" % (k[0], k[1], c, t / 1e6, t / c / 1e3, 100 * t / tot))
print("total %.3f ms over %d launches" % (tot / 1e6, sum(v[0] for v in agg.values())))
