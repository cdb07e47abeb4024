#!/bin/bash
mkdir -p gpurun_out
# skip the 100 weight-conversion launches: only our hot-path kernels match the regex anyway
timeout 900 ncu --set full --clock-control none --import-source on \
  -k regex:"projection_kernel|gemm_kernel|attention_kernel|layernorm_bf16_kernel|ln_pre_kernel" -c 10 \
  -o gpurun_out/prof_r01_v1 -f python scripts/prof_kernels.py > gpurun_out/prof.log 2>&1
echo "ncu exit $?"; tail -5 gpurun_out/prof.log; ls -la gpurun_out/
