#!/usr/bin/env python
"""Hot spots of an ncu source page: tools/ncu_hot.py report.ncu-rep [top]  (runs `ncu -i ... --page source --csv`).
Prints the instructions with the most stall samples, with the dominant stall reasons, and the totals per reason."""
import csv, subprocess, sys, collections
rep, top = sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 40
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout.splitlines()
rows = list(csv.reader(out))
hi = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
hdr = rows[hi]
col = {h: i for i, h in enumerate(hdr)}
stalls = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
recs = []
tot = collections.Counter()
for r in rows[hi + 1:]:
    if len(r) < len(hdr): continue
    n = int(r[col["# Samples"]] or 0)
    d = {s: int(r[col[s]] or 0) for s in stalls}
    for s, v in d.items(): tot[s] += v
    recs.append((n, r[col["Address"]], r[col["Source"]].strip(), d, int(r[col["Instructions Executed"]] or 0)))
total = sum(n for n, *_ in recs)
print("total samples", total, {k: v for k, v in tot.most_common(8)})
base = int(recs[0][1], 16)
for idx in sorted(range(len(recs)), key=lambda i: -recs[i][0])[:top]:
    n, addr, src, d, ex = recs[idx]
    why = ", ".join(f"{k[6:]}={v}" for k, v in sorted(d.items(), key=lambda kv: -kv[1])[:3] if v)
    print(f"{n:7d} {100 * n / total:5.1f}%  +{int(addr, 16) - base:06x} x{ex:<9d} {src[:70]:70s} {why}")
