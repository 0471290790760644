LRCN_TEST_PRECS=1 timeout 900 python -m pytest tests -x -q -m gpu 2>&1 | tail -3
LRCN_TEST_PRECS=0 timeout 900 python -m pytest tests -x -q -m gpu 2>&1 | tail -3
run() { timeout 300 python bench.py --no-cpu-baseline --no-beam > gpurun_out/bench_$1.log 2>&1; python -c "
import json,sys;d=json.loads(open('gpurun_out/bench_$1.log').read().strip().splitlines()[-1]);print('$1', d['value'], d['ms_per_step'], d['e2e']['value'], d['gpu_launches'])"; }
run pdl2
LRCN_PDL=1 run pdl1
LRCN_PDL=0 run pdl0
LRCN_NO_FUSED_SOFTMAX=1 run nofusedsmx
LRCN_NO_DUALB=1 run nodualb
run pdl2b
PROFILE_STEP=1 timeout 300 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches.csv python tools/prof_kernels.py none > gpurun_out/ncu_step.log 2>&1
python tools/summarize_launches.py gpurun_out/launches.csv
