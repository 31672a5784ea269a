cp streamkit_b200/csrc/libskgpu.so /tmp/keep.so
cp _variants/knobs/libskgpu.so streamkit_b200/csrc/libskgpu.so
A="--no-hub --no-router --no-s16-extra --no-capacity-check --steps 40 --warmup 5 --parity-sessions 0 --latency-ticks 0"
run() { env "$@" python bench.py $A 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
print('$*', 'ms_per_step', round(d['ms_per_step'],4), {k: round(v,4) for k,v in d.get('kernels_ms',{}).items()})
"; }
run X=1
run SKGPU_CHAIN_KB=1 SKGPU_CHAIN_STAGES=4
run SKGPU_CHAIN_KB=1 SKGPU_CHAIN_STAGES=3
run SKGPU_CHAIN_KB=1 SKGPU_CHAIN_STAGES=2
run SKGPU_CHAIN_KB=2 SKGPU_CHAIN_STAGES=3
run X=1
cp /tmp/keep.so streamkit_b200/csrc/libskgpu.so
