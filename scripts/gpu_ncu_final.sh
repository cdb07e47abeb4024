#!/bin/bash
mkdir -p gpurun_out
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches.csv \
   python bench.py --frames 8 --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/bench_under_ncu.json 2> gpurun_out/ncu_list.err
echo "ncu list exit $?"; wc -l gpurun_out/launches.csv
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"gemm2_kernel|attention_tc_kernel" -s 5 -c 10 -o gpurun_out/prof_vit_r01 -f \
   python scripts/prof_kernels.py > gpurun_out/prof_vit.log 2>&1; echo "ncu vit exit $?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"projection_kernel|gemm_kernel|ln_pre_kernel|head_kernel" -c 4 -o gpurun_out/prof_misc_r01 -f \
   python scripts/prof_kernels.py > gpurun_out/prof_misc.log 2>&1; echo "ncu misc exit $?"
ls -la gpurun_out/*.ncu-rep
