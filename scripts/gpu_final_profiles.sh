#!/bin/bash
# refresh of the per-config numbers under profiles/ with the final kernels
mkdir -p gpurun_out
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches.csv \
   python bench.py --frames 8 --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/bench_under_ncu.json 2> gpurun_out/ncu_list.err
echo "ncu list exit $?"; wc -l gpurun_out/launches.csv
timeout 300 ncu --set full --clock-control none --import-source on -k regex:"gemm2_kernel|ln_pre_kernel|head_kernel" -c 3 -o gpurun_out/prof_front_r01 -f \
   python scripts/prof_kernels.py > gpurun_out/prof_front.log 2>&1; echo "ncu front exit $?"
timeout 300 python scripts/bench_cfg1.py > gpurun_out/cfg1.jsonl 2> gpurun_out/cfg1.err; cat gpurun_out/cfg1.jsonl
timeout 600 python scripts/bench_encoder.py > gpurun_out/encoder_sweep.jsonl 2>gpurun_out/encoder_sweep.err; cat gpurun_out/encoder_sweep.jsonl
timeout 900 python bench.py --steps 2 --warmup 3 --frames 16 --clusters-per-frame 600 --n-max 16384 --no-cpu-baseline > gpurun_out/bench_cfg3_1gpu.json 2>gpurun_out/bench_cfg3.err; python -c "
import json; d=json.loads(open('gpurun_out/bench_cfg3_1gpu.json').read().strip().splitlines()[-1]); print('cfg3', d['value'], d['e2e']['value'], d['roofline_projection']['frac'], d['kernel_breakdown_rank0']['projection'])"
timeout 200 python scripts/bench_kernels.py > gpurun_out/kernels.json 2>/dev/null; cat gpurun_out/kernels.json
