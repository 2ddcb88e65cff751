"""Top CUDA source lines of a kernel by executed warp instructions / stall samples.
Input: `ncu -i X.ncu-rep --page source --csv --print-source cuda,sass -k regex:K > f.csv`."""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
n = int(sys.argv[2]) if len(sys.argv) > 2 else 40
hdr, out, cur = None, [], ""
for r in rows:
    if r and r[0] == "File Path": cur = r[1].split("/")[-1]; continue
    if r and r[0] == "Line No": hdr = r; continue
    if hdr and len(r) > 10: out.append((cur, r))
iI, iSm, iA = hdr.index("Instructions Executed"), hdr.index("# Samples"), hdr.index("Address")
iS = [i for i, h in enumerate(hdr) if h == "Source"][0]
def fl(x):
    try: return float(x)
    except Exception: return 0.0
src = [(f, r) for f, r in out if not r[iA].strip().lower().startswith("0x")]
ag = sorted(((fl(r[iI]), fl(r[iSm]), f, r[0], r[iS].strip()[:105]) for f, r in src if fl(r[iI]) > 0 or fl(r[iSm]) > 0), reverse=True)
ti, ts = sum(a[0] for a in ag) or 1, sum(a[1] for a in ag) or 1
print(f"# {ti:.3e} warp instructions, {ts:.0f} samples")
for a in ag[:n]:
    print(f"{a[0] / ti * 100:5.1f}% inst {a[1] / ts * 100:5.1f}% smp  {a[2]}:{a[3]:>4} {a[4]}")
