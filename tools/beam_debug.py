"""Host-loop accounting of lrcn_beam_search (LRCN_BEAM_DEBUG=1): how long the host loop ran and how much of it waited for the GPU."""
import os, sys
os.environ["LRCN_BEAM_DEBUG"] = "1"
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import lrcn_b200  # noqa
from lrcn_b200 import abi
import bench
for n in (32, 1024):
    for shaped in (False, True):
        r = bench.beam_leg(0, 0, 1, abi.PREC_BF16X3, lambda x: x, lambda: None, shaped, n_img=n, K=3)
        print(n, "shaped" if shaped else "worst", round(r["ms_per_batch"], 3), "ms", flush=True)
