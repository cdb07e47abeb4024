#!/bin/bash
mkdir -p gpurun_out
export PYTHONWARNINGS=ignore
timeout 600 compute-sanitizer --tool memcheck --error-exitcode 9 --print-limit 5 \
    python -m pytest tests/test_vit_gpu.py tests/test_e2e_gpu.py -x -q -m gpu -k "gemm or attention or tower or chunking" \
    > gpurun_out/sanitize_memcheck_tower.log 2>&1; echo "memcheck tower exit $?"
grep -E "ERROR SUMMARY|passed|failed|Error:" gpurun_out/sanitize_memcheck_tower.log | sort | uniq -c | head -8
timeout 600 compute-sanitizer --tool racecheck --error-exitcode 9 --print-limit 5 \
    python -m pytest tests/test_vit_gpu.py -x -q -m gpu -k "tower and bf16" \
    > gpurun_out/sanitize_racecheck_tower.log 2>&1; echo "racecheck tower exit $?"
grep -E "RACECHECK SUMMARY|passed|failed|Error:|hazard" gpurun_out/sanitize_racecheck_tower.log | sort | uniq -c | head -8
timeout 300 compute-sanitizer --tool synccheck --error-exitcode 9 --print-limit 5 \
    python -m pytest tests/test_vit_gpu.py tests/test_projection_gpu.py -x -q -m gpu -k "(tower and bf16 and ln) or large_clusters" \
    > gpurun_out/sanitize_synccheck.log 2>&1; echo "synccheck exit $?"
grep -E "ERROR SUMMARY|passed|failed|Error:" gpurun_out/sanitize_synccheck.log | sort | uniq -c | head -8
