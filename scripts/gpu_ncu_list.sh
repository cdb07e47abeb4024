#!/bin/bash
mkdir -p gpurun_out
# launch list of a reduced bench (8 frames, 1 step): per-launch device time, cold cache / serialised
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches.csv \
   python bench.py --frames 8 --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/bench_under_ncu.json 2> gpurun_out/ncu_list.err
echo "ncu list exit $?"; wc -l gpurun_out/launches.csv
# one --set full capture of the dominant kernel (fc GEMM, 2-CTA) for the DRAM traffic figure
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"gemm2_kernel" -s 2 -c 3 -o gpurun_out/prof_gemm2_r01 -f \
   python scripts/prof_kernels.py > gpurun_out/prof_gemm2.log 2>&1; echo "ncu gemm2 exit $?"
