"""Per-kernel device time and share of ONE step from an ncu launch list
(`ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file X.csv python bench.py ...`).
A step = the launches between two consecutive preprocess_kernel launches (cold-cache, serialised: compare shares)."""
import collections, csv, re, sys
rows = list(csv.reader(open(sys.argv[1])))
for i, r in enumerate(rows):
    if r and r[0] == "ID":
        hdr, start = r, i + 1
        break
idx = {h: i for i, h in enumerate(hdr)}
data = [r for r in rows[start:] if len(r) == len(hdr)]
names = [(r[idx["Kernel Name"]], float(r[idx["Metric Value"]].replace(",", "")), r[idx["Metric Unit"]]) for r in data]
pre = [i for i, n in enumerate(names) if "preprocess_kernel" in n[0]]
which = int(sys.argv[2]) if len(sys.argv) > 2 else -2
s, e = pre[which - 1], pre[which]
agg, tot = collections.OrderedDict(), 0.0
for n, v, u in names[s:e]:
    k = re.sub(r"\(.*", "", n).replace("void ", "")[:78]
    ms = v / 1e6 if u in ("ns", "nsecond") else (v / 1e3 if u in ("us", "usecond") else v)
    a = agg.setdefault(k, [0, 0.0]); a[0] += 1; a[1] += ms; tot += ms
print(f"# one step: {e - s} kernel launches, {tot:.3f} ms summed device time (serialised under ncu)")
for k, (c, v) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print(f"{v:8.3f} ms {v / tot * 100:5.1f} %  x{c:<3d} {k}")
