#!/bin/bash
# A/B of tuning knobs (library built with -DSKGPU_TUNING_KNOBS under _variants/knobs/): usage tools/ab_knobs.sh mix|chain
cp streamkit_b200/csrc/libskgpu.so /tmp/keep.so
cp _variants/knobs/libskgpu.so streamkit_b200/csrc/libskgpu.so
if [ "$1" = "mix" ]; then
  for t in 1 2 3 5 1; do echo "SKGPU_MIX_TPC=$t"; SKGPU_MIX_TPC=$t SK_ONLY=mix python tools/bench_kernels.py 2>/dev/null | python -c "
import json,sys
for l in sys.stdin:
    d=json.loads(l); print('   ', d['kernel'], round(d['ms']*1000,1), 'us', round(d['frac'],3))
"; done
else
  A="--no-hub --no-router --no-s16-extra --no-capacity-check --steps 40 --warmup 5 --parity-sessions 0 --latency-ticks 0"
  for kv in "X=1" "SKGPU_CHAIN_KB=1 SKGPU_CHAIN_STAGES=4" "SKGPU_CHAIN_KB=2 SKGPU_CHAIN_STAGES=3" "X=1"; do
    env $kv python bench.py $A 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
print('$kv', 'ms_per_step', round(d['ms_per_step'],4), {k: round(v,4) for k,v in d.get('kernels_ms',{}).items()})
"; done
fi
cp /tmp/keep.so streamkit_b200/csrc/libskgpu.so
