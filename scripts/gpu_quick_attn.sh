#!/bin/bash
mkdir -p gpurun_out
timeout 120 python -m pytest tests/test_vit_gpu.py -q -m gpu -k "attention" -x 2>&1 | grep -E "passed|failed|assert|Error" | head -5
ONLY=attention timeout 200 python scripts/bench_kernels.py
bash scripts/gpu_attn_trace.sh 2>&1 | tail -15
