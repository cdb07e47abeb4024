#!/bin/bash
# compute-sanitizer passes over the hand-written kernels on small inputs (slow: minutes)
mkdir -p gpurun_out
export PYTHONWARNINGS=ignore
for tool in memcheck racecheck; do
  timeout 420 compute-sanitizer --tool $tool --error-exitcode 9 --print-limit 5 \
    python -m pytest tests/test_projection_gpu.py tests/test_canonicalise_gpu.py -x -q -m gpu -k "rotation_modes or degenerate or large_clusters or canonicalis" \
    > gpurun_out/sanitize_${tool}_proj.log 2>&1; echo "$tool projection/canonicalise exit $?"
  grep -E "ERROR SUMMARY|RACECHECK SUMMARY|passed|failed|Error:|hazard" gpurun_out/sanitize_${tool}_proj.log | sort | uniq -c | head -8
done
timeout 420 compute-sanitizer --tool memcheck --error-exitcode 9 --print-limit 5 \
    python -m pytest tests/test_vit_gpu.py -x -q -m gpu -k "head_matches or vote or (attention and 1-)" \
    > gpurun_out/sanitize_memcheck_vit.log 2>&1; echo "memcheck vit exit $?"
grep -E "ERROR SUMMARY|passed|failed|Error:" gpurun_out/sanitize_memcheck_vit.log | sort | uniq -c | head -8
