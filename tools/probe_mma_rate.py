"""GPU probe: clocks per tcgen05.mma (bf16, SS mode, K = 16) as a function of the instruction shape."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import lrcn_b200  # noqa: F401,E402
from lrcn_b200 import abi  # noqa: E402

if __name__ == "__main__":
    cfg = abi.default_config(embed=64, hidden1=64, hidden2=64, vocab=100, max_batch=4, max_len=2, max_gen_rows=4, precision=1)
    with abi.Handle(cfg, hooks=True) as h:
        for (M, N) in ((128, 16), (128, 32), (128, 64), (64, 32), (64, 128), (128, 128)):
            for (ce, iss) in ((0, 1), (4, 1), (8, 1), (16, 1), (0, 2), (4, 2)):
                n = 512
                h.test_mma_rate(M, N, n, ce, iss)
                issue, total = h.test_mma_rate(M, N, n, ce, iss)
                print(f"M={M:3d} N={N:3d} commit_every={ce:2d} issuers={iss}: issue {issue / n:6.1f} clk/mma, complete {total / n:6.1f} clk/mma per issuer", flush=True)
