timeout 600 python tools/probe_gemm.py --quick --2cta > gpurun_out/probe_quick.log 2>&1
grep -c rel= gpurun_out/probe_quick.log; grep -E "FAIL|TIMEOUT" gpurun_out/probe_quick.log | head; awk '{for(i=1;i<=NF;i++) if ($i ~ /^rel=/) {split($i,a,"="); if (a[2]+0 > 2e-5) print}}' gpurun_out/probe_quick.log | head
echo "== TMA epilogue + stream-K" > gpurun_out/gemm_time.log; timeout 200 python tools/probe_gemm_time.py --dbg 0 >> gpurun_out/gemm_time.log 2>&1
cat gpurun_out/gemm_time.log
LRCN_TEST_PRECS=1 timeout 900 python -m pytest tests -x -q -m gpu 2>&1 | tail -5
LRCN_TEST_PRECS=0 timeout 900 python -m pytest tests -x -q -m gpu 2>&1 | tail -5
timeout 300 python bench.py --no-cpu-baseline --no-beam > gpurun_out/bench_pdl.log 2>&1; python -c "
import json;d=json.loads(open('gpurun_out/bench_pdl.log').read().strip().splitlines()[-1]);print('PDL on ', d['value'], d['ms_per_step'], d['e2e']['value'], d['gpu_launches'])"
LRCN_PDL=0 timeout 300 python bench.py --no-cpu-baseline --no-beam > gpurun_out/bench_nopdl.log 2>&1; python -c "
import json;d=json.loads(open('gpurun_out/bench_nopdl.log').read().strip().splitlines()[-1]);print('PDL off', d['value'], d['ms_per_step'], d['e2e']['value'], d['gpu_launches'])"
tail -3 gpurun_out/bench_pdl.log | cut -c1-600
