#!/bin/bash
# Round 2: projection sweep (quick or full) + ncu of the projection kernel on the Waymo mix.
mkdir -p gpurun_out
timeout 600 python scripts/bench_projection.py $1 > gpurun_out/proj_sweep.jsonl 2> gpurun_out/proj_sweep.err; echo "sweep exit $?"; cat gpurun_out/proj_sweep.jsonl; tail -3 gpurun_out/proj_sweep.err
