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
for c in config2 config3 config4 config4_sinc config4_down k1 s16 opus opus_s16; do [ -f gpurun_out/bench_${c}_$TAG.json ] && cp gpurun_out/bench_${c}_$TAG.json ${OUT}_bench_${c}.json; done
python tools/ncu_summary.py gpurun_out/prof_full_$TAG.ncu-rep $(ls gpurun_out/prof_nodes_$TAG.ncu-rep 2>/dev/null) > ${OUT}_ncu_full_summary.csv
python tools/ncu_lines.py gpurun_out/prof_full_$TAG.ncu-rep k_chain 65536 > ${OUT}_k_chain_hot_lines.txt 2>&1
python tools/ncu_lines.py gpurun_out/prof_full_$TAG.ncu-rep k_phase_chain 4096 > ${OUT}_k_phase_chain_hot_lines.txt 2>&1
[ -f gpurun_out/prof_nodes_$TAG.ncu-rep ] && python tools/ncu_lines.py gpurun_out/prof_nodes_$TAG.ncu-rep k_resample_sinc_tiled 16384 > ${OUT}_k_resample_sinc_hot_lines.txt 2>&1
[ -f gpurun_out/prof_nodes_$TAG.ncu-rep ] && python tools/ncu_lines.py gpurun_out/prof_nodes_$TAG.ncu-rep k_resample_prog 16384 > ${OUT}_k_resample_prog_hot_lines.txt 2>&1
cuobjdump -sass streamkit_b200/csrc/libskgpu.so | grep -oE "UBLKCP[.A-Z0-9]*|SYNCS[.A-Z0-9]*|FFMA2|FMUL2|FADD2|LDGSTS[.A-Z0-9]*" | sort | uniq -c > ${OUT}_sass_mnemonics.txt
ls -la profiles
