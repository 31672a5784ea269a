#!/bin/bash
# Produces the round's profiling artefacts under gpurun_out/ (copy the summaries you want judged into profiles/):
#   launches.csv      per-launch durations of one short bench run (ncu --metrics gpu__time_duration.sum)
#   prof_full.ncu-rep ncu --set full capture of k_phase_chain + k_chain (one launch each, steady state)
#   bench.json        the bench line of an unprofiled run (the only place numbers are taken from)
set -x
TAG=${1:-final}
python bench.py --steps 50 --warmup 5 > gpurun_out/bench_$TAG.json 2> gpurun_out/bench_$TAG.err
python bench.py --impl reference --steps 20 --warmup 3 > gpurun_out/bench_ref_$TAG.json 2>> gpurun_out/bench_$TAG.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_$TAG.csv python bench.py --steps 3 --warmup 3 > gpurun_out/launches_$TAG.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:"k_chain|k_phase_chain" -s 8 -c 2 -o gpurun_out/prof_full_$TAG -f python bench.py --steps 2 --warmup 3 > gpurun_out/ncu_full_$TAG.log 2>&1
python tools/bench_kernels.py > gpurun_out/kernels_$TAG.jsonl 2> gpurun_out/kernels_$TAG.err
tail -c 600 gpurun_out/bench_$TAG.json
