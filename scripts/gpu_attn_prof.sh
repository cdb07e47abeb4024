#!/bin/bash
mkdir -p gpurun_out
timeout 300 python scripts/bench_kernels.py > gpurun_out/kern_times.json 2>gpurun_out/kern_times.err; cat gpurun_out/kern_times.json; tail -2 gpurun_out/kern_times.err
ONLY=attention timeout 600 ncu --set full --clock-control none --import-source on -k regex:"attention_tc_kernel" -s 2 -c 1 -o gpurun_out/prof_attn_r01b -f python scripts/bench_kernels.py > gpurun_out/prof_attn.log 2>&1; echo "ncu exit $?"
