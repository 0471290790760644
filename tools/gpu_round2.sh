#!/bin/bash
# Round-2 GPU-box visit (1 GPU): parity tests, smoke, bench, ncu launch list + full captures of the LSTM sequence kernels.
mkdir -p gpurun_out
nvidia-smi > gpurun_out/nvidia_smi.txt 2>&1
if [ -z "$SKIP_TESTS" ]; then
  echo "== pytest -m gpu"; timeout 1500 python -m pytest tests -m gpu -q -x -p no:cacheprovider --durations=15 > gpurun_out/r02_pytest.log 2>&1; tail -25 gpurun_out/r02_pytest.log
  echo "== smoke"; timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r02_smoke.log 2>&1; tail -3 gpurun_out/r02_smoke.log
fi
if [ -z "$SKIP_BENCH" ]; then
  echo "== bench"; timeout 900 python bench.py --steps ${STEPS:-100} --warmup 5 > gpurun_out/r02_bench.log 2>&1; tail -c 6000 gpurun_out/r02_bench.log
fi
if [ -n "$EXTRA_BENCH" ]; then
  for wl in flickr8k_train_b64 coco_2f_train_b256; do
    timeout 600 python bench.py --workload $wl --steps 50 --warmup 5 --no-beam > gpurun_out/r02_bench_$wl.log 2>&1; grep '^{' gpurun_out/r02_bench_$wl.log | tail -1 | cut -c1-300
  done
fi
if [ -n "$BEAM_SWEEP" ]; then
  echo "== beam sweep (C5)"; timeout 600 python tools/beam_sweep.py > gpurun_out/r02_beam_sweep.md 2>gpurun_out/r02_beam_sweep.err; cat gpurun_out/r02_beam_sweep.md
fi
if [ -n "$SANITIZER" ]; then
  echo "== compute-sanitizer memcheck (subset: epoch staging, per-step LSTM kernels, beam search incl. compaction, top-K)"
  timeout 1200 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_gpu_parity.py -x -q -p no:cacheprovider \
      -k "epoch or (loss_and_gradients and 0-case0) or beam_search_matches or many_rows" > gpurun_out/r02_sanitizer_memcheck2.log 2>&1; echo "memcheck rc=$?"; tail -4 gpurun_out/r02_sanitizer_memcheck2.log
fi
if [ -z "$SKIP_NCU" ]; then
  echo "== ncu launch list"
  PROFILE_STEP=1 timeout 600 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r02_step_launches.csv \
      python tools/prof_kernels.py none > gpurun_out/r02_ncu_step.log 2>&1
  python tools/summarize_launches.py gpurun_out/r02_step_launches.csv > gpurun_out/r02_step_launches.md 2>&1; cat gpurun_out/r02_step_launches.md
  echo "== ncu full: LSTM sequence kernels"
  timeout 900 ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:"lstm_.*seq" -c 4 \
      -o gpurun_out/r02_prof_lstm -f python tools/prof_kernels.py lstm_fwd lstm_bwd > gpurun_out/r02_ncu_lstm.log 2>&1
  tail -5 gpurun_out/r02_ncu_lstm.log
  ncu -i gpurun_out/r02_prof_lstm.ncu-rep --page raw --csv > gpurun_out/r02_prof_lstm_raw.csv 2>/dev/null
  echo "== ncu full: the other hot kernels (one launch each)"
  timeout 900 ncu --profile-from-start off --set full --clock-control none -k regex:"gemm2_bf16x3|adam_kernel|softmax_ce|gather_embed" -c 8 \
      -o gpurun_out/r02_prof_other -f python tools/prof_kernels.py vocab_gemm adam softmax_ce gather > gpurun_out/r02_ncu_other.log 2>&1
  ncu -i gpurun_out/r02_prof_other.ncu-rep --page raw --csv > gpurun_out/r02_prof_other_raw.csv 2>/dev/null
  python - <<'PY'
import csv
a = list(csv.reader(open("gpurun_out/r02_prof_lstm_raw.csv")))
b = list(csv.reader(open("gpurun_out/r02_prof_other_raw.csv")))
if a[0] == b[0]:
    csv.writer(open("gpurun_out/r02_prof_all_raw.csv", "w")).writerows(a + b[2:])
else:  # different metric sets: align by header
    idx = [b[0].index(h) if h in b[0] else None for h in a[0]]
    csv.writer(open("gpurun_out/r02_prof_all_raw.csv", "w")).writerows(a + [[r[i] if i is not None else "0" for i in idx] for r in b[2:]])
PY
  python tools/ncu_to_traffic.py gpurun_out/r02_prof_all_raw.csv gpurun_out/r02_step_launches.csv flickr30k_train_b256 12 > gpurun_out/ncu_traffic.json; head -c 600 gpurun_out/ncu_traffic.json
  rm -f gpurun_out/r02_prof_other.ncu-rep
  ls -la gpurun_out/r02_* | head
fi
