"""One batched beam search (1024 images x beam 3, V=10000) between cuProfilerStart/Stop for an ncu launch list."""
import ctypes
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import lrcn_b200  # noqa: E402,F401
from lrcn_b200 import abi, synth  # noqa: E402

n_img, K, nword = 1024, 3, int(os.environ.get("NWORD", "5"))
cfg = abi.default_config(embed=512, hidden1=512, hidden2=512, vocab=10000, max_batch=8, max_len=2, max_gen_rows=n_img * K, precision=abi.PREC_BF16X3, use_graphs=0)
g = abi.Handle(cfg)
g.set_model(synth.initweights([512, 512], 10000, 512, seed=1))
ids = np.arange(1, n_img + 1, dtype=np.int64)
g.load_features(1, ids, synth.features(n_img, seed=6))
g.beam_search(1, ids, K, nword, want_logps=False)
cuda = ctypes.CDLL("libcuda.so.1")
cuda.cuProfilerStart()
t0 = time.perf_counter()
g.beam_search(1, ids, K, nword, want_logps=False)
print("ms", 1e3 * (time.perf_counter() - t0), "steps", nword + 1)
cuda.cuProfilerStop()
g.close()
