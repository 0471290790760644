"""Per-step timeline of the persistent forward LSTM kernel (CTA (0,0), layer 2): where the step-to-step latency goes."""
import os
import sys

import numpy as np

os.environ["LRCN_SEQ_TRACE"] = "1"
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import lrcn_b200  # noqa: E402,F401
from lrcn_b200 import abi, synth  # noqa: E402
import bench  # noqa: E402

w = bench.WORKLOADS["flickr30k_train_b256"]
cfg = abi.default_config(embed=w["E"], hidden1=w["H1"], hidden2=w["H2"], vocab=w["V"], max_batch=w["B"], max_len=28, max_gen_rows=8,
                         precision=abi.PREC_BF16X3, use_graphs=1)
h = abi.Handle(cfg)
h.set_model(synth.initweights([w["H1"], w["H2"]], w["V"], w["E"], seed=1))
h.load_features(0, np.arange(1, 1025, dtype=np.int64), synth.features(1024, seed=2))
l = 12
h.stage_batch(0, 0, synth.image_ids(w["B"], 1024), synth.tokens(l, w["B"], w["V"], zipf=True))
for i in range(4):
    h.train_step_staged(0, 0.0, i)
h.sync()
tr = h.get_trace(l + 1).astype(np.int64)
names = ["0 grid_wait done", "1 kb0 landed", "2 MMAs issued+commit", "3 epilogue: tfull seen", "4 all epilogue warps done (bar.sync)", "5 kb3 landed", "6 release done", "7 last kb landed"]
print("step | " + " | ".join(names) + "   (ns relative to this step's grid_wait)")
for t in range(1, l + 1):
    base = tr[t, 0]
    row = [int(tr[t, k] - base) if tr[t, k] else -1 for k in range(8)]
    nxt = tr[t + 1, 0] - base if t + 1 <= l and tr[t + 1, 0] else -1
    print(t, row, "next grid_wait at", nxt)
h.close()
