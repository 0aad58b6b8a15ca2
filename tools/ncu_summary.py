#!/usr/bin/env python
"""Summarise an .ncu-rep: headline raw metrics + SASS segments ranked by executed instructions."""
import collections
import csv
import subprocess
import sys

rep = sys.argv[1]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
hdr, units = rows[0], rows[1]
WANT = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "smsp__inst_executed.sum",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "launch__occupancy_limit_shared_mem",
        "launch__occupancy_limit_registers", "launch__shared_mem_per_block_dynamic", "launch__grid_size",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
        "lts__t_sectors_srcunit_tex_op_read.sum", "sm__cycles_elapsed.max",
        "smsp__average_warp_latency_issue_stalled_long_scoreboard.ratio", "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio"]
for r in rows[2:]:
    d = dict(zip(hdr, r))
    print("==", d.get("Kernel Name"))
    for w in WANT:
        if w in d:
            print(f"{w}: {d[w]} {units[hdr.index(w)]}")
    st = []
    for k, v in d.items():
        if "issue_stalled" in k and k.endswith("_per_warp_active.pct"):
            try:
                st.append((float(v), k.replace("smsp__average_warps_issue_stalled_", "").replace("_per_warp_active.pct", "")))
            except ValueError:
                pass
    print("stalls(% warp active):", ", ".join(f"{k}={v:.1f}" for v, k in sorted(st, reverse=True)[:8]))
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(src.splitlines()))
hdr, data = rows[1], rows[2:]
ie, sc, smp = hdr.index("Instructions Executed"), hdr.index("Source"), hdr.index("# Samples")
tot = sum(int(r[ie]) for r in data)
totsmp = sum(int(r[smp]) for r in data)
segs, cur = [], None
for i, r in enumerate(data):
    c = int(r[ie])
    if cur and (0.5 * cur["c"] <= c <= 2 * cur["c"]):
        cur["n"] += 1; cur["sum"] += c; cur["smp"] += int(r[smp]); cur["end"] = i
    else:
        cur = {"start": i, "end": i, "c": max(c, 1), "n": 1, "sum": c, "smp": int(r[smp])}; segs.append(cur)
segs.sort(key=lambda s: -s["sum"])
print(f"total warp instructions {tot}, stall samples {totsmp}")
for s in segs[: int(sys.argv[2]) if len(sys.argv) > 2 else 10]:
    ops = []
    for i in range(s["start"], s["end"] + 1):
        t = data[i][sc].split()
        ops.append(t[1] if t[0].startswith("@") else t[0])
    print(f"idx {s['start']}-{s['end']} n={s['n']} exec/inst~{s['sum']//s['n']} share={100*s['sum']/tot:.1f}% samples={100*s['smp']/max(totsmp,1):.1f}%",
          collections.Counter(ops).most_common(8))
