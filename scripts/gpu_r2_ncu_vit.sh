#!/bin/bash
# Round 2: launch list of a short bench run (kernel SHARES of the step) + full ncu capture of one layer's
# GEMM / attention kernels at 4096 images.
mkdir -p gpurun_out
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r02_launches.csv \
   python bench.py --frames 8 --steps 1 --warmup 3 --no-cpu-baseline --no-extras > gpurun_out/bench_under_ncu.json 2> gpurun_out/ncu_list.err
echo "ncu list exit $?"; wc -l gpurun_out/r02_launches.csv
IMAGES=4096 timeout 900 ncu --set full --clock-control none --import-source on -k regex:"gemm2_kernel|attention_tc_kernel" -s 1 -c 5 -o gpurun_out/r02_vit -f \
   python scripts/prof_kernels.py > gpurun_out/prof_vit.log 2>&1; echo "ncu vit exit $?"; tail -2 gpurun_out/prof_vit.log
ls -la gpurun_out/*.ncu-rep
