#!/bin/bash
mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none --import-source on \
  -k regex:"attention_tc_kernel|gemm2_kernel|layernorm_bf16_kernel" -c 6 \
  -o gpurun_out/prof_r01_v3 -f python scripts/prof_kernels.py > gpurun_out/prof.log 2>&1
echo "ncu exit $?"; tail -3 gpurun_out/prof.log
