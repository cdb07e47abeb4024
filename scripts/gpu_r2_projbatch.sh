#!/bin/bash
# Round 2: in-step projection time with the whole batch projected first (default) vs one chunk at a time.
mkdir -p gpurun_out
for pb in ${PBS:-default 4096}; do
  if [ $pb = default ]; then unset VG_PROJ_BATCH; else export VG_PROJ_BATCH=$pb; fi
  timeout 600 python bench.py --steps 2 --warmup 3 --no-extras --no-cpu-baseline > gpurun_out/bench_pb_$pb.json 2> gpurun_out/bench_pb_$pb.err
  echo "pb=$pb exit $?"; tail -2 gpurun_out/bench_pb_$pb.err
  python - <<PY
import json
d=json.load(open("gpurun_out/bench_pb_$pb.json"))
print("value", d["value"], "e2e", d["e2e"]["value"], "ms", d["ms_per_step"])
print("roofline_projection", {k: d["roofline_projection"][k] for k in ("achieved","frac","avg_launch_ms","launches","share_of_step")})
print("alone", d["roofline_projection_alone"]["frac"], "gemm frac", d["roofline"]["frac"])
print({k: round(v["ms_per_step"],1) for k,v in d["kernel_breakdown_rank0"].items()})
print(d["clocks"])
PY
done
