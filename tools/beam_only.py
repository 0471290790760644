import sys, os
sys.path.insert(0, "/root/repo")
import lrcn_b200
from lrcn_b200 import abi
import bench
for shaped in (True, False):
    r = bench.beam_leg(0, 0, 1, abi.PREC_BF16X3, lambda x: x, lambda: None, shaped=shaped)
    print("shaped" if shaped else "worst", round(r["value"]), round(r["ms_per_batch"], 3), flush=True)
