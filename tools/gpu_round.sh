#!/bin/bash
# One GPU-box visit: probe kernels, parity tests (fp32 first, then tcgen05), smoke, bench.  Logs -> gpurun_out/
mkdir -p gpurun_out
nvidia-smi > gpurun_out/nvidia_smi.txt 2>&1
[ -n "$SKIP_PROBE" ] || { echo "== probe"; timeout 600 python tools/probe_gemm.py --quick > gpurun_out/probe.log 2>&1; grep -c "rel=" gpurun_out/probe.log; grep -E "FAIL|TIMEOUT|rel=[0-9.]+e-0[0-4]" gpurun_out/probe.log | head -20; }
echo "== pytest fp32"; LRCN_TEST_PRECS=0 timeout 900 python -m pytest tests -m gpu -q -p no:cacheprovider > gpurun_out/pytest_fp32.log 2>&1; tail -6 gpurun_out/pytest_fp32.log
echo "== pytest bf16x3"; LRCN_TEST_PRECS=1 timeout 900 python -m pytest tests -m gpu -q -p no:cacheprovider > gpurun_out/pytest_bf16.log 2>&1; tail -12 gpurun_out/pytest_bf16.log
echo "== smoke"; timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; tail -3 gpurun_out/smoke.log
[ -n "$SKIP_FP32_BENCH" ] || { echo "== bench fp32"; timeout 600 python bench.py --steps 10 --warmup 3 --precision fp32 --no-cpu-baseline > gpurun_out/bench_fp32.log 2>&1; tail -2 gpurun_out/bench_fp32.log; }
echo "== bench bf16x3"; timeout 600 python bench.py --steps 20 --warmup 3 > gpurun_out/bench_bf16.log 2>&1; tail -2 gpurun_out/bench_bf16.log
echo "== profile"; timeout 600 bash tools/profile.sh > gpurun_out/profile.log 2>&1; tail -3 gpurun_out/profile.log
