#!/bin/bash
# Round 2: A/B/A of the packed-pair GEMM epilogues (default build) against the scalar form (-DVG_EPI_SCALAR) on one box.
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_vit_gpu.py -q -m gpu -x 2>&1 | tail -1
for v in new old new2; do
  if [ $v = old ]; then export VG_LIB_PATH=vilgod_b200/lib/libvilgod_b200_epi0.so; else unset VG_LIB_PATH; fi
  timeout 600 python bench.py --steps 2 --warmup 3 --no-extras --no-cpu-baseline > gpurun_out/bench_epi_$v.json 2> gpurun_out/bench_epi_$v.err
  python - <<PY
import json
d=json.load(open("gpurun_out/bench_epi_$v.json"))
print("$v", "value %.1f" % d["value"], "gemm frac %.4f" % d["roofline"]["frac"], {k: round(x["ms_per_step"],1) for k,x in d["kernel_breakdown_rank0"].items() if k.startswith("gemm") or k=="attention"}, d["clocks"]["sm_mhz"])
PY
done
