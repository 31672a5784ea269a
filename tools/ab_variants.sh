#!/bin/bash
# A/B helper: builds of libskgpu.so under _variants/<name>/ are swapped in one after the other and benched on the same box.
# usage (on the GPU box): tools/ab_variants.sh "<bench args>" name1 name2 ...
ARGS=$1; shift
cp streamkit_b200/csrc/libskgpu.so /tmp/libskgpu_keep.so
for v in "$@"; do
  cp _variants/$v/libskgpu.so streamkit_b200/csrc/libskgpu.so
  for rep in 1 2; do
    python bench.py $ARGS 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
print('$v', 'ms_per_step', round(d['ms_per_step'],4), 'kernels', {k: round(v,4) for k,v in d.get('kernels_ms',{}).items()}, 'frac', round(d['roofline']['frac'],3))
"
  done
done
cp /tmp/libskgpu_keep.so streamkit_b200/csrc/libskgpu.so
