#!/usr/bin/env python
"""Condenses `ncu --set full` reports into one CSV (header row, unit row, one row per captured launch) with the columns
the round's analysis cites. usage: tools/ncu_summary.py a.ncu-rep [b.ncu-rep ...] > profiles/rN_ncu_full_summary.csv"""
import csv
import subprocess
import sys

KEEP = ["ID", "Kernel Name", "launch__grid_size", "launch__block_size", "launch__registers_per_thread", "launch__shared_mem_per_block_dynamic",
        "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem", "gpu__time_duration.sum", "dram__bytes_read.sum",
        "dram__bytes_write.sum", "dram__throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "smsp__inst_executed.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_alu.sum.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fma.sum.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_fp64.sum.pct_of_peak_sustained_active", "sm__inst_executed_pipe_lsu.sum.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_xu.sum.pct_of_peak_sustained_active", "sm__inst_executed_pipe_tma.sum.pct_of_peak_sustained_active",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "lts__t_sector_hit_rate.pct", "smsp__pcsamp_warps_issue_stalled_long_scoreboard",
        "smsp__pcsamp_warps_issue_stalled_short_scoreboard", "smsp__pcsamp_warps_issue_stalled_wait", "smsp__pcsamp_warps_issue_stalled_not_selected",
        "smsp__pcsamp_warps_issue_stalled_selected", "smsp__pcsamp_warps_issue_stalled_branch_resolving",
        "smsp__pcsamp_warps_issue_stalled_math_pipe_throttle", "smsp__pcsamp_warps_issue_stalled_no_instructions",
        "smsp__pcsamp_warps_issue_stalled_barrier", "smsp__pcsamp_warps_issue_stalled_mio_throttle", "smsp__pcsamp_warps_issue_stalled_lg_throttle"]


def main():
    w = csv.writer(sys.stdout)
    units = None
    out = []
    for rep in sys.argv[1:]:
        txt = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
        rows = [r for r in csv.reader(txt.splitlines()) if r]
        hdr, u = rows[0], rows[1]
        if units is None:
            units = dict(zip(hdr, u))
        for r in rows[2:]:
            d = dict(zip(hdr, r))
            uu = dict(zip(hdr, u))
            scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "ns": 1e-9, "us": 1e-6, "usecond": 1e-6, "ms": 1e-3, "msecond": 1e-3, "s": 1.0, "second": 1.0, "nsecond": 1e-9}
            for k in ("dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__time_duration.sum"):      # one unit per column across reports
                if k in d and uu.get(k) != units.get(k) and uu.get(k) in scale and units.get(k) in scale:
                    d[k] = "%g" % (float(d[k].replace(",", "")) * scale[uu[k]] / scale[units[k]])
            d["Kernel Name"] = d.get("Kernel Name", "")[:90]
            out.append(d)
    w.writerow(KEEP)
    w.writerow([units.get(k, "") for k in KEEP])
    for d in out:
        w.writerow([d.get(k, "") for k in KEEP])


if __name__ == "__main__":
    main()
