#!/usr/bin/env python
"""tools/ncu_hot.py <report.ncu-rep> <kernel regex> [launch index] -- hottest SASS instructions (stall samples) of one
captured launch, with the dominant stall reason, from `ncu --page source --csv`."""
import csv, io, subprocess, sys
rep, rx = sys.argv[1], sys.argv[2]
nth = sys.argv[3] if len(sys.argv) > 3 else "1"
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-id", f"::regex:{rx}:{nth}"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
print(rows[0][1])
h = rows[1]; idx = {x: i for i, x in enumerate(h)}
stalls = [c for c in h if c.startswith("stall_") and "Not Issued" not in c]
body = [r for r in rows[2:] if len(r) == len(h) and r[0].startswith("0x")]
sc = idx["# Samples"]
tot = sum(float(r[sc] or 0) for r in body)
inst = sum(float(r[idx["Instructions Executed"]] or 0) for r in body)
print(f"samples {tot:.0f}  warp instructions {inst:.0f}")
agg = {s: sum(float(r[idx[s]] or 0) for r in body) for s in stalls}
print("stall mix:", "  ".join(f"{k[6:]}={v / tot * 100:.0f}%" for k, v in sorted(agg.items(), key=lambda kv: -kv[1])[:8]))
for i, r in sorted(enumerate(body), key=lambda ir: -float(ir[1][sc] or 0))[:int(sys.argv[4]) if len(sys.argv) > 4 else 30]:
    top = max(stalls, key=lambda s: float(r[idx[s]] or 0))
    print(f"{float(r[sc]) / tot * 100:5.1f}%  #{i:4d} {r[idx['Source']].strip()[:70]:70s} {top[6:]}")
