#!/bin/bash
# profiling helper: device-resident tick time of the fused chain for different staging-ring depths / debug modes
for s in 2 3 4; do
  echo -n "stages $s: "
  SKGPU_CHAIN_STAGES=$s python bench.py --steps 30 --warmup 5 2>/dev/null | python -c 'import json,sys;d=json.loads(sys.stdin.read());print(d["ms_per_step"],d["kernels_ms"])'
done
echo -n "nocompute: "
SKGPU_CHAIN_DEBUG=1 python bench.py --steps 30 --warmup 5 2>/dev/null | python -c 'import json,sys;d=json.loads(sys.stdin.read());print(d["ms_per_step"],d["kernels_ms"])'
