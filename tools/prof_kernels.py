"""Run one training step of the bench workload, then launch each hot kernel family once between
cuProfilerStart/Stop so that `ncu --profile-from-start off --set full` captures exactly those launches."""
import ctypes
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import lrcn_b200  # noqa: E402,F401
from lrcn_b200 import abi, synth  # noqa: E402
import bench  # noqa: E402

w = bench.WORKLOADS[os.environ.get("WORKLOAD", "flickr30k_train_b256")]
cfg = abi.default_config(embed=w["E"], hidden1=w["H1"], hidden2=w["H2"], vocab=w["V"], max_batch=w["B"], max_len=28, max_gen_rows=8,
                         precision=abi.PREC_BF16X3, use_graphs=0)
h = abi.Handle(cfg)
h.set_model(synth.initweights([w["H1"], w["H2"]], w["V"], w["E"], seed=1))
h.load_features(0, np.arange(1, 1025, dtype=np.int64), synth.features(1024, seed=2))
l = 12
img = synth.image_ids(w["B"], 1024)
tok = synth.tokens(l, w["B"], w["V"], zipf=True)
h.stage_batch(0, 0, img, tok)
PDROP = float(os.environ.get("PDROP", "0.4"))  # lrcn.jl:227
for i in range(2):
    h.train_step_staged(0, PDROP, i)
h.sync()
cuda = ctypes.CDLL("libcuda.so.1")
cuda.cuProfilerStart()
names = sys.argv[1:] or ["vocab_gemm", "gate_gemm", "adam", "softmax_ce", "gather", "lstm_fwd", "lstm_bwd"]
for nm in names:
    if nm == "none":
        continue
    ms, by, fl = h.time_kernel(nm, 1)
    print(nm, "ms", ms, "GB/s", by / ms / 1e6, "TFLOP/s", fl / ms / 1e9, flush=True)
if os.environ.get("PROFILE_STEP"):
    h.train_step_staged(0, PDROP, 5)
    h.sync()
cuda.cuProfilerStop()
h.close()
