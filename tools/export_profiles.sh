#!/bin/bash
# Copies the judged summaries of a profile_round.sh run from gpurun_out/ into profiles/ (tracked).
# usage: tools/export_profiles.sh <tag-in-gpurun_out> <name-prefix-in-profiles>
set -e
TAG=$1; OUT=profiles/$2
cp gpurun_out/bench_$TAG.json ${OUT}_bench.json
cp gpurun_out/bench_ref_$TAG.json ${OUT}_bench_reference.json
cp gpurun_out/kernels_$TAG.jsonl ${OUT}_kernels_configs234.jsonl
grep -v "^==" gpurun_out/launches_$TAG.csv | python -c '
import csv,sys
rows=list(csv.reader(sys.stdin)); hdr=rows[0]
keep=[hdr.index(k) for k in ("ID","Kernel Name","Block Size","Grid Size","Metric Name","Metric Unit","Metric Value")]
w=csv.writer(sys.stdout)
for r in rows:
    if len(r)==len(hdr): w.writerow([r[i][:60] for i in keep])
' > ${OUT}_launches.csv
ncu -i gpurun_out/prof_full_$TAG.ncu-rep --page raw --csv 2>/dev/null | python -c '
import csv,sys
rows=list(csv.reader(sys.stdin)); hdr=rows[0]
keep=["ID","Kernel Name","launch__grid_size","launch__block_size","launch__registers_per_thread","launch__shared_mem_per_block_dynamic","launch__occupancy_limit_registers","launch__occupancy_limit_shared_mem","gpu__time_duration.sum","dram__bytes_read.sum","dram__bytes_write.sum","dram__throughput.avg.pct_of_peak_sustained_elapsed","sm__throughput.avg.pct_of_peak_sustained_elapsed","smsp__inst_executed.sum","smsp__issue_active.avg.pct_of_peak_sustained_active","sm__warps_active.avg.pct_of_peak_sustained_active","sm__inst_executed_pipe_alu.sum.pct_of_peak_sustained_active","sm__inst_executed_pipe_fma.sum.pct_of_peak_sustained_active","sm__inst_executed_pipe_fp64.sum.pct_of_peak_sustained_active","sm__inst_executed_pipe_lsu.sum.pct_of_peak_sustained_active","sm__inst_executed_pipe_xu.sum.pct_of_peak_sustained_active","sm__inst_executed_pipe_tma.sum.pct_of_peak_sustained_active","l1tex__data_pipe_lsu_wavefronts_mem_shared.sum","lts__t_sector_hit_rate.pct","smsp__pcsamp_warps_issue_stalled_long_scoreboard","smsp__pcsamp_warps_issue_stalled_short_scoreboard","smsp__pcsamp_warps_issue_stalled_wait","smsp__pcsamp_warps_issue_stalled_not_selected","smsp__pcsamp_warps_issue_stalled_selected","smsp__pcsamp_warps_issue_stalled_branch_resolving","smsp__pcsamp_warps_issue_stalled_math_pipe_throttle","smsp__pcsamp_warps_issue_stalled_no_instructions"]
idx=[hdr.index(k) for k in keep if k in hdr]
w=csv.writer(sys.stdout)
for r in rows: w.writerow([r[i] for i in idx])
' > ${OUT}_ncu_full_summary.csv
python tools/ncu_lines.py gpurun_out/prof_full_$TAG.ncu-rep k_chain 65536 > ${OUT}_k_chain_hot_lines.txt 2>&1
python tools/ncu_lines.py gpurun_out/prof_full_$TAG.ncu-rep k_phase_chain 4096 > ${OUT}_k_phase_chain_hot_lines.txt 2>&1
cuobjdump -sass streamkit_b200/csrc/libskgpu.so | grep -oE "UBLKCP[.A-Z0-9]*|SYNCS[.A-Z0-9]*|FFMA2|FMUL2|LDGSTS[.A-Z0-9]*" | sort | uniq -c > ${OUT}_sass_mnemonics.txt
ls -la profiles
