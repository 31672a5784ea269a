#!/bin/bash
# compute-sanitizer passes over the GPU parity suite (run on a B200 box: gpurun -- 'bash tools/sanitize.sh')
cd "$(dirname "$0")/.."
SEL='not 16384 and not 4096_sessions and not long_run'
for tool in memcheck synccheck initcheck; do
  echo "== $tool"
  timeout 1200 compute-sanitizer --tool $tool --error-exitcode 3 python -m pytest tests -m gpu -x -q -k "$SEL" 2>&1 | grep -E "passed|failed|ERROR SUMMARY|Error:|=========     at" | head -20
done
