#!/bin/bash
# Round 2: the driver's own multi-GPU launch line for N ranks on one box.
N=${1:-2}
mkdir -p gpurun_out
nvidia-smi -L | head -8
timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 \
    bench.py --gpus $N --steps ${STEPS:-2} --warmup 3 > gpurun_out/bench_${N}gpu.json 2> gpurun_out/bench_${N}gpu.err
echo "bench $N gpus exit $?"; tail -3 gpurun_out/bench_${N}gpu.err
python - <<PY
import json
d = json.loads(open("gpurun_out/bench_${N}gpu.json").read().strip().splitlines()[-1])
for k in ("value", "n_gpus", "ms_per_step", "e2e", "cfg3", "strong_scaling", "clocks"):
    print(k, d.get(k))
PY
