#!/bin/bash
mkdir -p gpurun_out
timeout 300 python scripts/diag_accuracy.py > gpurun_out/diag.log 2>&1; tail -12 gpurun_out/diag.log
timeout 900 python bench.py --steps 2 --warmup 3 > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench exit $?"; tail -c 6000 gpurun_out/bench.json; tail -5 gpurun_out/bench.err
