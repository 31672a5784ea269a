#!/bin/bash
# Produces the round's profiling artefacts under gpurun_out/ (copy the summaries you want judged into profiles/):
#   launches.csv      per-launch durations of one short bench run (ncu --metrics gpu__time_duration.sum)
#   prof_full.ncu-rep ncu --set full capture of k_phase_chain + k_chain (one launch each, steady state)
#   prof_nodes.ncu-rep ncu --set full capture of k_convert / k_mix / k_resample_prog / k_resample_sinc (tools/bench_kernels.py)
#   bench*.json       the bench lines of unprofiled runs (the only place numbers are taken from)
set -x
# SKIP_NCU_FULL=1: bench lines and the launch list only (the two --set full reports together exceed what gpurun copies back
# when the repo already holds them: capture them in a call of their own)
TAG=${1:-final}
python bench.py --steps 50 --warmup 5 > gpurun_out/bench_$TAG.json 2> gpurun_out/bench_$TAG.err
python bench.py --impl reference --steps 20 --warmup 3 > gpurun_out/bench_ref_$TAG.json 2>> gpurun_out/bench_$TAG.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_$TAG.csv python bench.py --steps 3 --warmup 3 > gpurun_out/launches_$TAG.log 2>&1
[ -z "$SKIP_NCU_FULL" ] && ncu --set full --clock-control none --import-source on -k regex:"k_chain|k_phase_chain" -s 8 -c 2 -o gpurun_out/prof_full_$TAG -f python bench.py --steps 2 --warmup 3 > gpurun_out/ncu_full_$TAG.log 2>&1
python tools/bench_kernels.py > gpurun_out/kernels_$TAG.jsonl 2> gpurun_out/kernels_$TAG.err
# the standalone node kernels (configs 2/3/4 + the sinc mode): second launch of every workload
[ -z "$SKIP_NCU_FULL" ] && SK_PROFILE=1 ncu --set full --clock-control none -k regex:"k_convert|k_mix|k_resample_prog|k_resample_sinc" -o gpurun_out/prof_nodes_$TAG -f python tools/bench_kernels.py > gpurun_out/ncu_nodes_$TAG.log 2>&1
for c in 2 3 4; do python bench.py --config $c --steps 30 --warmup 5 > gpurun_out/bench_config${c}_$TAG.json 2>> gpurun_out/bench_$TAG.err; done
python bench.py --config 4 --sinc --steps 30 --warmup 5 > gpurun_out/bench_config4_sinc_$TAG.json 2>> gpurun_out/bench_$TAG.err
python bench.py --config 4 --rs-down --steps 30 --warmup 5 > gpurun_out/bench_config4_down_$TAG.json 2>> gpurun_out/bench_$TAG.err
# workload variants of the chain: one input per session, s16 ingest, Opus-decoder shaped (48 kHz mono inputs, bypass)
V="--steps 30 --warmup 5 --no-hub --no-router --no-s16-extra --no-capacity-check"
python bench.py $V --k 1 > gpurun_out/bench_k1_$TAG.json 2>> gpurun_out/bench_$TAG.err
python bench.py $V --s16-in > gpurun_out/bench_s16_$TAG.json 2>> gpurun_out/bench_$TAG.err
python bench.py $V --in-rate 48000 --channels 1 --k 3 > gpurun_out/bench_opus_$TAG.json 2>> gpurun_out/bench_$TAG.err
python bench.py $V --in-rate 48000 --channels 1 --k 3 --s16-in > gpurun_out/bench_opus_s16_$TAG.json 2>> gpurun_out/bench_$TAG.err
tail -c 600 gpurun_out/bench_$TAG.json
