#!/bin/bash
# Round 2 (third session), last call: compute-sanitizer over the reworked fast projection kernel, then the whole GPU suite + smoke.
mkdir -p gpurun_out
export PYTHONWARNINGS=ignore
for tool in memcheck racecheck; do
  timeout 100 compute-sanitizer --tool $tool --error-exitcode 9 --print-limit 5 \
    python -m pytest tests/test_projection_gpu.py -x -q -m gpu -k "fast_and_general" \
    > gpurun_out/r02_sanitize_${tool}_fast2.log 2>&1; echo "$tool exit $?"
  grep -E "ERROR SUMMARY|RACECHECK SUMMARY|passed|failed|Error:|hazard" gpurun_out/r02_sanitize_${tool}_fast2.log | sort | uniq -c | head -8
done
( timeout 150 python -m pytest tests -q -m gpu 2>&1 | grep -E "passed|failed|^E " | tail -4; timeout 60 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1 ) | tee gpurun_out/r02_gpu_tests_last.txt
