#!/bin/bash
# Round 2: GPU test groups, each under its own timeout, logs into gpurun_out/.
mkdir -p gpurun_out
nvidia-smi -L
run() { name=$1; shift; echo "=== $name"; timeout ${TMO:-600} "$@" > gpurun_out/$name.log 2>&1; echo "exit $?"; tail -n ${TAILN:-25} gpurun_out/$name.log; }
run t_projection python -m pytest tests/test_projection_gpu.py -q -m gpu
run t_canon python -m pytest tests/test_canonicalise_gpu.py -q -m gpu
run t_vit python -m pytest tests/test_vit_gpu.py -q -m gpu -s
run t_e2e python -m pytest tests/test_e2e_gpu.py -q -m gpu -s
run t_smoke python -c "import __graft_entry__ as g; g.smoke()"
