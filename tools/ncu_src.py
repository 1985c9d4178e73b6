#!/usr/bin/env python
"""Summarise an `ncu --page source --csv` dump: per-instruction executed counts, stall samples, smem conflicts."""
import csv
import sys

path = sys.argv[1]
top = int(sys.argv[2]) if len(sys.argv) > 2 else 0
rows = list(csv.reader(open(path)))
hdr = rows[1]
ix = {h: i for i, h in enumerate(hdr)}
recs = []
for r in rows[2:]:
    if len(r) < len(hdr):
        continue
    recs.append(dict(src=r[ix["Source"]].strip(), n=int(r[ix["Instructions Executed"]]), samples=int(r[ix["# Samples"]]),
                     wf=int(r[ix["L1 Wavefronts Shared"]]), wfi=int(r[ix["L1 Wavefronts Shared Ideal"]]),
                     stalls={k: int(r[ix[k]]) for k in hdr if k.startswith("stall_") and "Not Issued" not in k}))
tot = sum(x["n"] for x in recs)
ts = sum(x["samples"] for x in recs)
print(f"total warp-instructions {tot}, samples {ts}")
if top:
    for i, x in sorted(enumerate(recs), key=lambda t: -t[1]["samples"])[:top]:
        st = ", ".join(f"{k[6:]}={v}" for k, v in sorted(x["stalls"].items(), key=lambda kv: -kv[1])[:3] if v)
        print(f"{i:4d} n={x['n']:>10d} samp={x['samples']:>5d} wf={x['wf']}/{x['wfi']}  {x['src'][:70]:70s} {st}")
else:
    for i, x in enumerate(recs):
        st = ", ".join(f"{k[6:]}={v}" for k, v in sorted(x["stalls"].items(), key=lambda kv: -kv[1])[:2] if v)
        print(f"{i:4d} n={x['n']:>10d} samp={x['samples']:>5d} wf={x['wf']}/{x['wfi']}  {x['src'][:70]:70s} {st}")
