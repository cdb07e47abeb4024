#!/bin/bash
# Round 2 (third session): one full ncu capture (source-level) of the final fast projection kernel on the Waymo mix.
mkdir -p gpurun_out
timeout 300 ncu --set full --clock-control none --import-source on -k regex:projection_fast_kernel -s 3 -c 1 \
    -f -o gpurun_out/${1:-r02_proj_fast_v6} python scripts/bench_projection.py one > gpurun_out/ncu_proj_fast.log 2>&1
echo "ncu exit $?"; tail -3 gpurun_out/ncu_proj_fast.log; ls -la gpurun_out/*fast*.ncu-rep
