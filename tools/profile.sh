#!/bin/bash
# ncu evidence for one bench step (1 GPU only): launch list with device times, then a full capture of the top kernels.
mkdir -p gpurun_out
ncu --metrics gpu__time_duration.sum --clock-control none -s ${NCU_SKIP:-900} -c ${NCU_COUNT:-260} --csv --log-file gpurun_out/launches.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_bench.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:gemm_bf16x3 -s 60 -c 2 -o gpurun_out/prof_gemm -f \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_gemm.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:adam_kernel -c 1 -o gpurun_out/prof_adam -f \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_adam.log 2>&1
ls -la gpurun_out
