#!/bin/bash
# Round 2: compute-sanitizer over the kernels rewritten this round (projection: stamp / dense / R = 224;
# GEMM: residual planes, LNF variants, patch) on small inputs.
mkdir -p gpurun_out
export PYTHONWARNINGS=ignore
for tool in memcheck racecheck; do
  timeout 600 compute-sanitizer --tool $tool --error-exitcode 9 --print-limit 5 \
    python -m pytest tests/test_projection_gpu.py -x -q -m gpu -k "stamp_and_dense or r224 or degenerate or rotation_modes or reciprocal" \
    > gpurun_out/r02_sanitize_${tool}_proj.log 2>&1; echo "$tool projection exit $?"
  grep -E "ERROR SUMMARY|RACECHECK SUMMARY|passed|failed|Error:|hazard" gpurun_out/r02_sanitize_${tool}_proj.log | sort | uniq -c | head -8
done
timeout 600 compute-sanitizer --tool memcheck --error-exitcode 9 --print-limit 5 \
    python -m pytest tests/test_vit_gpu.py tests/test_e2e_gpu.py -x -q -m gpu -k "(residual and 988) or (layernorm_folded and 988) or patch_embedding or (tower and f16 and ln) or empty_frames or write_back" \
    > gpurun_out/r02_sanitize_memcheck_vit.log 2>&1; echo "memcheck vit exit $?"
grep -E "ERROR SUMMARY|passed|failed|Error:" gpurun_out/r02_sanitize_memcheck_vit.log | sort | uniq -c | head -8
timeout 600 compute-sanitizer --tool racecheck --error-exitcode 9 --print-limit 5 \
    python -m pytest tests/test_vit_gpu.py -x -q -m gpu -k "(residual and 988 and f16) or (tower and f16 and ln)" \
    > gpurun_out/r02_sanitize_racecheck_vit.log 2>&1; echo "racecheck vit exit $?"
grep -E "RACECHECK SUMMARY|passed|failed|Error:|hazard" gpurun_out/r02_sanitize_racecheck_vit.log | sort | uniq -c | head -8
