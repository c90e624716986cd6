"""Executed-instruction histogram by opcode from `ncu --page source --csv` output (one kernel)."""
import csv, sys, collections
rows = list(csv.reader(open(sys.argv[1])))
hi = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
hdr = rows[hi]
si, ei, pi = hdr.index("Source"), hdr.index("Instructions Executed"), hdr.index("Predicated-On Thread Instructions Executed")
px = float(sys.argv[2]) if len(sys.argv) > 2 else 0
seen, hist, tot = set(), collections.Counter(), 0
for r in rows[hi + 1:]:
    if len(r) <= pi or r[0] in seen or not r[0].startswith("0x"):
        continue  # the first launch only (the file repeats per captured launch)
    seen.add(r[0])
    src = r[si].strip()
    parts = src.split()
    op = parts[1] if parts[0].startswith("@") else parts[0]
    op = op.split(".")[0].rstrip(";")
    n = int(float(r[ei] or 0))
    hist[op] += n
    tot += n
print("total warp-instructions executed:", tot, ("= %.2f per pixel-lane" % (tot * 32 / px) if px else ""))
for op, n in hist.most_common(30):
    print(f"{op:12s} {n:12d} {100*n/tot:5.1f}%" + (f"  {n*32/px:6.2f}/px" if px else ""))
