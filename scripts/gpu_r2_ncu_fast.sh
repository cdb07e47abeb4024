#!/bin/bash
# Round 2: parity tests of the projection, then one full ncu capture (source-level) of the fast kernel on the Waymo mix.
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_projection_gpu.py -q -m gpu -x 2>&1 | grep -v Warning | grep -E "^E |passed|failed|Error" | head -30
timeout 900 ncu --set full --clock-control none --import-source on -k regex:projection_fast_kernel -s 3 -c 1 \
    -f -o gpurun_out/${1:-r02_proj_fast} python scripts/bench_projection.py one > gpurun_out/ncu_proj_fast.log 2>&1
echo "ncu exit $?"; tail -3 gpurun_out/ncu_proj_fast.log; ls -la gpurun_out/*fast*.ncu-rep
