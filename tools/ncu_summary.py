#!/usr/bin/env python
"""Turns an .ncu-rep into the text summary committed under profiles/: headline metrics of the first kernel in the
report plus the instructions with the most stall samples.  Usage: ncu_summary.py report.ncu-rep > profiles/x.txt"""
import csv
import io
import subprocess
import sys

rep = sys.argv[1]
top = int(sys.argv[2]) if len(sys.argv) > 2 else 25
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units, vals = rows[0], rows[1], rows[2]
d = {h: (u, v) for h, u, v in zip(hdr, units, vals)}
keys = [
    "Kernel Name", "gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
    "launch__shared_mem_per_block_dynamic", "launch__occupancy_limit_shared_mem", "launch__occupancy_limit_registers",
    "sm__cycles_elapsed.avg", "sm__cycles_elapsed.avg.per_second", "sm__cycles_active.avg",
    "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__bytes_read.sum.per_second",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
    "smsp__inst_executed.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "sm__warps_active.avg.pct_of_peak_sustained_active",
    "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed",
    "l1tex__data_bank_conflicts_pipe_lsu_mem_shared_op_ld.sum",
    "l1tex__data_pipe_lsu_wavefronts_mem_shared_op_ld.sum",
]
print(f"# {rep}")
for k in keys:
    if k in d:
        print(f"{k:75s} {d[k][1]:>18s} {d[k][0]}")
print("\n# warp stall reasons, cycles per issued instruction")
for h in hdr:
    if h.startswith("smsp__average_warps_issue_stalled_") and h.endswith("_per_issue_active.ratio"):
        v = float(d[h][1])
        if v >= 0.02:
            print(f"  {h[len('smsp__average_warps_issue_stalled_'):-len('_per_issue_active.ratio')]:28s} {v:.3f}")
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
h2 = rows[1]
ix = {h: i for i, h in enumerate(h2)}
recs = []
for r in rows[2:]:
    if len(r) < len(h2):
        continue
    recs.append((int(r[ix["# Samples"]]), int(r[ix["Instructions Executed"]]), r[ix["Source"]].strip()))
tot_i = sum(r[1] for r in recs)
tot_s = sum(r[0] for r in recs)
ff = sum(r[1] for r in recs if r[2].startswith("FFMA2"))
lds = sum(r[1] for r in recs if r[2].startswith("LDS"))
print(f"\n# instruction mix: {tot_i} warp instructions, FFMA2 {ff} ({100.0 * ff / max(tot_i, 1):.1f} %), LDS {lds}, "
      f"other {tot_i - ff - lds}; issue cycles by the 2-per-FFMA2 model: {tot_i + ff}")
print(f"# top {top} instructions by stall samples (of {tot_s})")
for s, n, t in sorted(recs, reverse=True)[:top]:
    print(f"  {s:6d} samples  executed {n:>10d}  {t}")
