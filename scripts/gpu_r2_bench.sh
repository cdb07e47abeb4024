#!/bin/bash
# Round 2: bench line of the default (fp16-operand) build and of the bf16 alternative on the same box.
mkdir -p gpurun_out
timeout 900 python bench.py --steps 3 --warmup 3 > gpurun_out/bench_f16.json 2> gpurun_out/bench_f16.err; echo "bench f16 exit $?"; tail -c 5000 gpurun_out/bench_f16.json; tail -3 gpurun_out/bench_f16.err
timeout 600 python bench.py --steps 3 --warmup 3 --operand-dtype bf16 --no-cpu-baseline > gpurun_out/bench_bf16.json 2> gpurun_out/bench_bf16.err; echo "bench bf16 exit $?"; tail -c 1500 gpurun_out/bench_bf16.json
