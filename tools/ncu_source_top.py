"""Top CUDA source lines of a kernel by stall samples.
Input: `ncu -i X.ncu-rep --page source --csv --print-source cuda,sass --kernel-name K > f.csv`."""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
n = int(sys.argv[2]) if len(sys.argv) > 2 else 30
cur_file, hdr, idx, out = None, None, None, []
for r in rows:
    if not r:
        continue
    if r[0] == "File Name":
        cur_file = r[1].split("/")[-1]
        continue
    if r[0] == "Line No":
        hdr = r
        idx = {h: i for i, h in enumerate(hdr)}
        continue
    if hdr and r[0].isdigit() and "# Samples" in idx and len(r) > idx["# Samples"]:
        def f(k):
            try:
                return float(r[idx[k]])
            except Exception:
                return 0.0
        out.append((f("# Samples"), f("Instructions Executed"), cur_file, r[0], r[1].strip()))
ts = sum(o[0] for o in out) or 1
ti = sum(o[1] for o in out) or 1
for s, i, fl, ln, src in sorted(out, key=lambda o: -o[0])[:n]:
    print(f"{s / ts * 100:5.1f}% smp {i / ti * 100:5.1f}% inst  {fl}:{ln}  {src[:120]}")
