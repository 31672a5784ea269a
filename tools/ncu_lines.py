#!/usr/bin/env python
"""Summarise an .ncu-rep: per kernel headline metrics and the hottest CUDA source lines (instructions / samples).
usage: ncu_lines.py report.ncu-rep [kernel-substring] [units-per-launch]"""
import csv, subprocess, sys, io
rep = sys.argv[1]
want = sys.argv[2] if len(sys.argv) > 2 else ""
units = float(sys.argv[3]) if len(sys.argv) > 3 else 1.0
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr = rows[0]
keys = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "smsp__inst_executed.sum",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "launch__registers_per_thread", "sm__inst_executed_pipe_xu.sum.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_lsu.sum.pct_of_peak_sustained_active", "sm__inst_executed_pipe_alu.sum.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_fma.sum.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fp64.sum.pct_of_peak_sustained_active",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "dram__throughput.avg.pct_of_peak_sustained_elapsed"]
ids = []
for r in rows[2:]:
    d = dict(zip(hdr, r))
    if want and want not in d["Kernel Name"]:
        continue
    ids.append(d["ID"])
    print("==", d["ID"], d["Kernel Name"][:60])
    for k in keys:
        if k in d:
            print("   ", k, d[k])
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--print-source", "cuda,sass", "--csv"], capture_output=True, text=True).stdout
cur = None
kern = None
out = {}
for r in csv.reader(io.StringIO(src)):
    if len(r) == 2 and r[0] == "Kernel Name": kern = r[1]; continue
    if len(r) == 2 and r[0] == "Function Name": kern = r[1]; continue
    if len(r) == 2 and r[0] == "File Path": cur = r[1].split("/")[-1]; continue
    if len(r) > 8 and r[0] not in ("", "Line No"):
        try:
            out.setdefault(kern, []).append((cur, int(r[0]), r[1], int(r[6]), int(r[7]) / units))
        except ValueError:
            pass
for kern, o in out.items():
    if want and want not in (kern or ""):
        continue
    print("=====", (kern or "")[:70], "samples", sum(x[3] for x in o), "inst/unit %.1f" % sum(x[4] for x in o))
    for x in sorted(o, key=lambda x: -x[4])[:int(40)]:
        print(f"{x[0]}:{x[1]:4d} smp={x[3]:6d} inst={x[4]:8.1f}  {x[2][:95]}")
