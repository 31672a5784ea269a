#!/bin/bash
cd /root/repo
timeout 600 compute-sanitizer --tool memcheck --error-exitcode 3 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "fused_chain_other_packet or full_chain_bit_exact or resampler_streaming_parity" 2>&1 | tail -15
echo "memcheck rc=$?"
timeout 600 compute-sanitizer --tool racecheck --error-exitcode 3 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "full_chain_bit_exact and True and 44100 and not 6-2" 2>&1 | tail -15
echo "racecheck rc=$?"
