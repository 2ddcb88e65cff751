"""Per-SASS-instruction execution profile of one kernel.
Input: `ncu -i X.ncu-rep --page source --csv --print-source cuda,sass -k regex:K > f.csv`; prints instructions in address order
with their share of executed warp instructions and stall samples."""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
hdr, out = None, []
for r in rows:
    if r and r[0] == "Line No":
        hdr = r
        continue
    if hdr and len(r) > 10:
        out.append(r)
iS = [i for i, h in enumerate(hdr) if h == "Source"]
iI, iSm, iA, iT = hdr.index("Instructions Executed"), hdr.index("# Samples"), hdr.index("Address"), hdr.index("Avg. Threads Executed")
def fl(x):
    try: return float(x)
    except Exception: return 0.0
sass = [r for r in out if r[iA].strip().lower().startswith("0x") or r[iA].strip().isdigit()]
seen, uniq = set(), []
for r in sass:
    if r[iA] in seen: continue
    seen.add(r[iA]); uniq.append(r)
uniq.sort(key=lambda r: int(r[iA], 16) if r[iA].startswith("0x") else int(r[iA]))
tot = sum(fl(r[iI]) for r in uniq) or 1
ts = sum(fl(r[iSm]) for r in uniq) or 1
print(f"# {len(uniq)} SASS instructions, {tot:.3e} warp instructions executed, {ts:.0f} samples")
thr = float(sys.argv[2]) if len(sys.argv) > 2 else 0.0
for r in uniq:
    c = fl(r[iI]) / tot * 100
    if c >= thr:
        print(f"{c:5.2f}% inst {fl(r[iSm]) / ts * 100:5.2f}% smp thr {fl(r[iT]):4.1f} L{r[0]:>4}  {r[iS[1]][:100]}")
