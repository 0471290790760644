#!/bin/bash
# ncu evidence (1 GPU only).  (1) launch list with device times of ONE full training step;
# (2) --set full captures of the hot kernel families launched alone.
mkdir -p gpurun_out
PROFILE_STEP=1 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches.csv \
    python tools/prof_kernels.py none > gpurun_out/ncu_step.log 2>&1
ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:"gemm|adam_kernel|softmax_ce|gather_embed" \
    -o gpurun_out/prof_kernels -f python tools/prof_kernels.py > gpurun_out/ncu_kernels.log 2>&1
tail -8 gpurun_out/ncu_kernels.log
ls -la gpurun_out | head -30
PROFILE_STEP=1 ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:"lstm_.*_seq" -c 2 \
    -o gpurun_out/prof_lstm_seq -f python tools/prof_kernels.py none > gpurun_out/ncu_lstm_seq.log 2>&1
