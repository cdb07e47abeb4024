#!/bin/bash
mkdir -p gpurun_out
nvidia-smi -L
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus 2 --steps 2 --warmup 3 --frames 16 > gpurun_out/bench_2gpu.json 2> gpurun_out/bench_2gpu.err; echo "2gpu exit $?"; tail -c 1500 gpurun_out/bench_2gpu.json; tail -5 gpurun_out/bench_2gpu.err
timeout 300 python bench.py --gpus 1 --steps 2 --warmup 3 --frames 16 --no-cpu-baseline > gpurun_out/bench_1gpu_16f.json 2>/dev/null; python -c "
import json
for f in ('gpurun_out/bench_1gpu_16f.json','gpurun_out/bench_2gpu.json'):
    d=json.loads(open(f).read().strip().splitlines()[-1]); print(f, d['n_gpus'], d['value'], d['e2e']['value'], d['ms_per_step'])
"
