#!/bin/bash
# compute-sanitizer passes over the GPU parity suite (run on a B200 box: gpurun -- 'bash tools/sanitize.sh [tools...]').
# The full-size, many-stream and long-running tests are left out (the sanitizer slows kernels 10-100x); every kernel and
# every input kind is still covered by the small cases.
cd "$(dirname "$0")/.."
SEL='not fullsize and not 16384 and not 4096_sessions and not long_run and not many_streams and not racing and not router and not stats_and_bounded'
TOOLS=${@:-memcheck synccheck initcheck}
for tool in $TOOLS; do
  echo "== $tool"
  timeout 1500 compute-sanitizer --tool $tool --error-exitcode 3 python -m pytest tests -m gpu -x -q -k "$SEL" 2>&1 | grep -E "passed|failed|ERROR SUMMARY|Error:|=========     at|Invalid|Uninitialized|hazard" | head -30
done
