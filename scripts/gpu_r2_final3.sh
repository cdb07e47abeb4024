#!/bin/bash
# Round 2 (third session): evidence set from the final tree -- GPU tests + smoke, the default bench arm, projection sweep.
mkdir -p gpurun_out
( python -m pytest tests -q -m gpu 2>&1 | grep -v "Warning\|wrap(\|k3 = \|^$\|Docs:\|warnings summary" | tail -8; python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2 ) > gpurun_out/r02_gpu_tests.txt 2>&1
cat gpurun_out/r02_gpu_tests.txt
timeout 420 python bench.py --steps 3 --warmup 3 > gpurun_out/r02_bench_default.json 2> gpurun_out/bench_default.err; echo "bench exit $?"; tail -3 gpurun_out/bench_default.err; cut -c1-600 gpurun_out/r02_bench_default.json
timeout 200 python scripts/bench_projection.py > gpurun_out/r02_projection_sweep.jsonl 2> gpurun_out/proj_sweep.err; echo "sweep exit $?"; cut -c1-200 gpurun_out/r02_projection_sweep.jsonl
