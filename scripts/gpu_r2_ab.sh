#!/bin/bash
# A/B of an environment switch on the same box: bench without extras, kernel breakdown side by side.
# usage: gpu_r2_ab.sh VG_SWITCH
mkdir -p gpurun_out
SW=$1
timeout 600 python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-extras > gpurun_out/ab_off.json 2>/dev/null; echo "off exit $?"
env $SW=1 timeout 600 python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-extras > gpurun_out/ab_on.json 2>/dev/null; echo "on exit $?"
timeout 600 python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-extras > gpurun_out/ab_off2.json 2>/dev/null; echo "off2 exit $?"
python - <<PY
import json
for f in ("ab_off", "ab_on", "ab_off2"):
    d = json.loads(open(f"gpurun_out/{f}.json").read().strip().splitlines()[-1])
    kb = {k: round(v["ms_per_step"], 1) for k, v in d["kernel_breakdown_rank0"].items() if v["ms_per_step"] > 1}
    print(f, round(d["value"], 1), d["clocks"]["sm_mhz"], d["parity"]["top1_raw"], kb)
PY
