#!/bin/bash
# Round 2 (second session): evidence set from the final tree -- GPU tests, both bench arms, projection sweep,
# ncu launch list of a short bench run.
mkdir -p gpurun_out
( python -m pytest tests -q -m gpu 2>&1 | grep -v "Warning\|wrap(\|k3 = \|^$\|Docs:\|warnings summary" | tail -8; python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2 ) > gpurun_out/r02_gpu_tests.txt 2>&1
cat gpurun_out/r02_gpu_tests.txt
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r02_bench_reference_arm.json 2> gpurun_out/ref_arm.err; echo "reference arm exit $?"; cut -c1-300 gpurun_out/r02_bench_reference_arm.json
timeout 1500 python bench.py --steps 3 --warmup 3 > gpurun_out/r02_bench_default.json 2> gpurun_out/bench_default.err; echo "bench exit $?"; tail -3 gpurun_out/bench_default.err
timeout 600 python scripts/bench_projection.py > gpurun_out/r02_projection_sweep.jsonl 2> gpurun_out/proj_sweep.err; echo "sweep exit $?"; cut -c1-230 gpurun_out/r02_projection_sweep.jsonl
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r02_ncu_launch_list.csv \
   python bench.py --frames 8 --steps 1 --warmup 3 --no-cpu-baseline --no-extras > gpurun_out/bench_under_ncu.json 2> gpurun_out/ncu_list.err
echo "ncu list exit $?"; wc -l gpurun_out/r02_ncu_launch_list.csv
