#!/bin/bash
# Round 2: one full ncu capture (source-level stalls, DRAM bytes) of the projection kernel on the Waymo mix.
mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none --import-source on -k regex:projection_kernel -s 3 -c 1 \
    -f -o gpurun_out/${1:-r02_proj} python scripts/bench_projection.py one > gpurun_out/ncu_proj.log 2>&1
echo "ncu exit $?"; tail -5 gpurun_out/ncu_proj.log; ls -la gpurun_out/*.ncu-rep
