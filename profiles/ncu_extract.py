#!/usr/bin/env python
"""Summarise an .ncu-rep: key raw metrics + hottest SASS lines with stall reasons.
usage: python profiles/ncu_extract.py gpurun_out/x.ncu-rep [min_share]"""
import csv, subprocess, sys, io

rep = sys.argv[1]
share = float(sys.argv[2]) if len(sys.argv) > 2 else 0.01
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units = rows[0], rows[1]
want = ["Kernel Name", "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "launch__registers_per_thread", "launch__grid_size", "launch__block_size", "launch__occupancy_limit", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed", "sm__inst_executed_pipe_tensor_subpipe_dmma.avg.pct_of_peak_sustained_active",
        "smsp__pipe_tensor_subpipe_dmma_cycles_active.avg", "smsp__cycles_active.avg", "sm__cycles_elapsed.max", "sm__issue_active.avg.pct_of_peak_sustained_elapsed",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared", "lts__t_bytes.sum", "lts__t_sector_hit_rate.pct", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "l1tex__t_bytes.sum", "smsp__average_warps_issue_stalled", "dram__throughput", "lts__throughput", "l1tex__throughput"]
for r in rows[2:]:
    print("=" * 100)
    for h, u, v in zip(hdr, units, r):
        if any(h == w or (h.startswith(w) and ("stalled" in w or "bank" in w or "throughput" in w or w.endswith("sum") and h == w) ) for w in want):
            if "stalled" in h and float(v or 0) < 0.2:
                continue
            print(f"{h:90s} {u:12s} {v}")
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
sections, cur = [], None          # one section per profiled launch
for r in rows:
    if r and r[0] == "Kernel Name":
        cur = {"name": r[1], "hdr": None, "data": []}
        sections.append(cur)
    elif r and r[0] == "Address" and cur is not None:
        cur["hdr"] = r
    elif cur is not None and cur["hdr"] is not None and len(r) >= len(cur["hdr"]):
        cur["data"].append(r)
for sec in sections:
    hdr, data = sec["hdr"], sec["data"]
    ix = {h: i for i, h in enumerate(hdr)}
    tot = sum(int(r[ix["# Samples"]] or 0) for r in data)
    print("-" * 100)
    print(sec["name"][:90], "total samples", tot)
    keys = [k for k in hdr if k.startswith("stall_") and "Not Issued" not in k]
    agg = {k: sum(int(r[ix[k]] or 0) for r in data) for k in keys}
    print("stall totals:", {k: v for k, v in sorted(agg.items(), key=lambda kv: -kv[1]) if v > tot * 0.005})
    for r in data:
        s = int(r[ix["# Samples"]] or 0)
        if s >= tot * share:
            st = {k[6:]: int(r[ix[k]] or 0) for k in keys if int(r[ix[k]] or 0) > s * 0.1}
            print(r[ix["Address"]][-5:], r[ix["Source"]][:64].ljust(64), s, r[ix["Instructions Executed"]], st)
