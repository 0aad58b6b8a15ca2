#!/usr/bin/env python
"""Print pipe utilisation, stall-reason samples and headline metrics of an .ncu-rep (first kernel)."""
import csv, subprocess, sys
raw = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
h, r = rows[0], rows[2]
d = dict(zip(h, r))
def f(k):
    try: return float(d[k].replace(",", ""))
    except Exception: return None
for k in ["gpu__time_duration.sum", "sm__cycles_elapsed.max", "smsp__inst_executed.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active",
          "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
          "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
          "l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed", "lts__t_sectors_srcunit_tex_op_read.sum",
          "launch__registers_per_thread", "launch__shared_mem_per_block_dynamic", "sm__warps_active.avg.pct_of_peak_sustained_active"]:
    print(k, d.get(k))
print("-- pipes (pct of peak, active)")
for k in sorted(d):
    if k.startswith("sm__inst_executed_pipe_") and k.endswith(".avg.pct_of_peak_sustained_active") or "pipe_xu_realtime" in k or k.startswith("sm__pipe_") and k.endswith("avg.pct_of_peak_sustained_active"):
        v = f(k)
        if v and v > 1: print(f"  {k}: {v:.1f}")
print("-- pc sampling stalls")
st = [(f(k), k.replace("smsp__pcsamp_warps_issue_stalled_", "")) for k in d if k.startswith("smsp__pcsamp_warps_issue_stalled_") and not k.endswith("_not_issued")]
tot = sum(v for v, _ in st if v)
for v, k in sorted([s for s in st if s[0]], reverse=True)[:12]:
    print(f"  {k}: {v:.0f} ({100*v/tot:.1f}%)")
