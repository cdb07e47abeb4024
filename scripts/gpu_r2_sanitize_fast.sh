#!/bin/bash
# Round 2: compute-sanitizer (memcheck, racecheck) over the fast projection kernel and the list-mode hand-over.
mkdir -p gpurun_out
export PYTHONWARNINGS=ignore
for tool in memcheck racecheck; do
  timeout 900 compute-sanitizer --tool $tool --error-exitcode 9 --print-limit 5 \
    python -m pytest tests/test_projection_gpu.py -x -q -m gpu -k "fast_and_general or degenerate or rotation_modes or adversarial or r224" \
    > gpurun_out/r02_sanitize_${tool}_fast.log 2>&1; echo "$tool exit $?"
  grep -E "ERROR SUMMARY|RACECHECK SUMMARY|passed|failed|Error:|hazard" gpurun_out/r02_sanitize_${tool}_fast.log | sort | uniq -c | head -8
done
