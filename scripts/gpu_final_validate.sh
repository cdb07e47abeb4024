#!/bin/bash
# what the driver runs at round end: GPU tests, smoke, default bench, reference arm
mkdir -p gpurun_out
timeout 900 python -m pytest tests/ -x -q -m gpu > gpurun_out/t_all.log 2>&1; echo "pytest -m gpu exit $?"; grep -v Warning gpurun_out/t_all.log | grep -E "passed|failed|^E " | head
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -3
timeout 900 python bench.py > gpurun_out/bench_default.json 2> gpurun_out/bench_default.err; echo "bench exit $?"
timeout 600 python bench.py --impl reference > gpurun_out/bench_reference.json 2> gpurun_out/bench_reference.err; echo "ref exit $?"
python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench_default.json').read().strip().splitlines()[-1])
print({k:d[k] for k in ('metric','value','unit','n_gpus','steps','warmup','ms_per_step','scaling','vs_baseline','dtype','gpu_launches')})
print('e2e', d['e2e']); print('roofline', d['roofline']); print('cpu', d['cpu_baseline']); print(d['clocks'])
r=json.loads(open('gpurun_out/bench_reference.json').read().strip().splitlines()[-1])
print('ref', r['value'], r['cpu_baseline'])
PY
