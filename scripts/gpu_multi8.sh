#!/bin/bash
# the driver's own N-GPU launch line (task statement), N from $1 (default 8)
N=${1:-8}
mkdir -p gpurun_out
nvidia-smi -L | head -8
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29531 bench.py --gpus $N --steps 2 --warmup 3 > gpurun_out/bench_${N}gpu.json 2> gpurun_out/bench_${N}gpu.err; echo "${N}gpu exit $?"
tail -c 2500 gpurun_out/bench_${N}gpu.json; grep -v Warning gpurun_out/bench_${N}gpu.err | tail -5
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29532 bench.py --impl reference --gpus $N --steps 1 --warmup 0 > gpurun_out/bench_${N}gpu_ref.json 2> gpurun_out/bench_${N}gpu_ref.err; echo "ref arm exit $?"; tail -c 600 gpurun_out/bench_${N}gpu_ref.json
