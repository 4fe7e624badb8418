"""Top SASS instructions by stall samples from `ncu -i rep --page source --csv` (reads the csv path)."""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
hdr = rows[1]
ix = {h: i for i, h in enumerate(hdr)}
stall_cols = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
data = []
tot = 0
for r in rows[2:]:
    try:
        s = int(r[ix["# Samples"]])
    except Exception:
        continue
    tot += s
    data.append((s, r))
top = int(sys.argv[2]) if len(sys.argv) > 2 else 30
agg = {c: 0 for c in stall_cols}
for s, r in data:
    for c in stall_cols:
        try: agg[c] += int(r[ix[c]])
        except Exception: pass
print("total samples", tot)
print("by reason:", ", ".join("%s=%.1f%%" % (c[6:], 100 * v / max(tot, 1)) for c, v in sorted(agg.items(), key=lambda kv: -kv[1])[:8]))
for s, r in sorted(data, key=lambda x: -x[0])[:top]:
    reasons = sorted(((int(r[ix[c]] or 0), c[6:]) for c in stall_cols), reverse=True)[:2]
    print("%6.2f%% %8s exec  %-70s %s" % (100 * s / tot, r[ix["Instructions Executed"]], r[ix["Source"]].strip()[:70],
                                         " ".join("%s:%d" % (n, v) for v, n in reasons if v)))
