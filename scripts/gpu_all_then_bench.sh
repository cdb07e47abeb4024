#!/bin/bash
mkdir -p gpurun_out
timeout 240 python -m pytest tests/test_vit_gpu.py -q -m gpu -k "gemm or attention" -x > gpurun_out/t_kern.log 2>&1; rc=$?
echo "gemm+attention tests exit $rc"; grep -v Warning gpurun_out/t_kern.log | grep -E "assert|Error|passed|failed|^E " | head -20
if [ $rc -ne 0 ]; then exit 0; fi
timeout 600 python -m pytest tests/test_vit_gpu.py tests/test_e2e_gpu.py tests/test_projection_gpu.py -q -m gpu -k "not gemm and not attention" -s > gpurun_out/t_rest.log 2>&1; echo "rest exit $?"; grep -v Warning gpurun_out/t_rest.log | grep -E "assert|agreement|passed|failed|^E " | head
timeout 900 python bench.py --steps 2 --warmup 3 ${BENCH_ARGS:---no-cpu-baseline} > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench exit $?"; python - <<'PY'
import json
try:
    d=json.loads(open('gpurun_out/bench.json').read().strip().splitlines()[-1])
    print('value', d['value'], 'e2e', d['e2e']['value'], 'ms/step', d['ms_per_step'])
    print('roofline', d['roofline']['achieved'], d['roofline']['frac'], 'vit frac', d['vit_tensor_frac_of_peak'])
    print('proj', d['roofline_projection']['achieved'], d['roofline_projection']['frac'])
    for k,v in d['kernel_breakdown_rank0'].items(): print(f"  {k:12s} {v['ms_per_step']:9.1f} ms")
    print(d['clocks'])
except Exception as e:
    print('parse fail', e); print(open('gpurun_out/bench.err').read()[-2000:])
PY
