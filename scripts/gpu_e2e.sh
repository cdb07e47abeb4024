#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_e2e_gpu.py -q -m gpu -s > gpurun_out/t_e2e.log 2>&1; echo "exit $?"; grep -v Warning gpurun_out/t_e2e.log | grep -E "assert|Error|agreement|passed|failed|^E " | head -30
