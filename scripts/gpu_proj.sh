#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_projection_gpu.py tests/test_e2e_gpu.py -q -m gpu -x -k "not cfg1_frame" > gpurun_out/t_proj.log 2>&1; echo "proj+graph tests exit $?"
grep -v Warning gpurun_out/t_proj.log | grep -E "assert|Error|passed|failed|^E " | head -20
timeout 600 python scripts/bench_projection.py > gpurun_out/proj_sweep.jsonl 2> gpurun_out/proj_sweep.err; echo "sweep exit $?"
python - <<'PY'
import json
for l in open('gpurun_out/proj_sweep.jsonl'):
    d=json.loads(l); print(d['views'],d['points_per_cluster'],d['clusters'],round(d['us_per_image'],2),round(d['frac_of_measured_hbm_peak'],3))
PY
