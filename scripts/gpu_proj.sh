#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_projection_gpu.py -q -m gpu -x > gpurun_out/t_projection.log 2>&1; rc=$?
echo "projection tests exit $rc"; grep -E "passed|failed|Error|error|assert" gpurun_out/t_projection.log | tail -15
timeout 300 python scripts/bench_projection.py > gpurun_out/proj_sweep.jsonl 2> gpurun_out/proj_sweep.err; echo "sweep exit $?"; cat gpurun_out/proj_sweep.jsonl; tail -3 gpurun_out/proj_sweep.err
timeout 300 python -m pytest tests/test_e2e_gpu.py -q -m gpu -s 2>&1 | grep -E "agreement|passed|failed"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:projection_kernel -c 1 -o gpurun_out/prof_proj_v2 -f python scripts/prof_kernels.py > gpurun_out/prof2.log 2>&1; echo "ncu exit $?"
