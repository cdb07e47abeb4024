#!/bin/bash
mkdir -p gpurun_out
timeout 600 python scripts/bench_encoder.py > gpurun_out/encoder_sweep.jsonl 2>gpurun_out/encoder_sweep.err; cat gpurun_out/encoder_sweep.jsonl; tail -2 gpurun_out/encoder_sweep.err
# cfg3: Argoverse-2 shaped (600 clusters / frame, up to 16k points), 16 frames on one GPU
timeout 900 python bench.py --steps 2 --warmup 3 --frames 16 --clusters-per-frame 600 --n-max 16384 --no-cpu-baseline > gpurun_out/bench_cfg3_1gpu.json 2>gpurun_out/bench_cfg3.err; python -c "
import json; d=json.loads(open('gpurun_out/bench_cfg3_1gpu.json').read().strip().splitlines()[-1]); print('cfg3', d['value'], d['e2e']['value'], d['roofline_projection']['frac'], d['kernel_breakdown_rank0']['projection'])"
timeout 300 python scripts/bench_projection.py > gpurun_out/proj_sweep.jsonl 2>/dev/null; cat gpurun_out/proj_sweep.jsonl | python -c "
import sys,json
for l in sys.stdin:
    d=json.loads(l); print(d['views'], d['points_per_cluster'], round(d['us_per_image'],4), round(d['frac_of_measured_hbm_peak'],4))"
