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
        n = 512
        for (M, N) in ((128, 64), (128, 128)):
            for flags, what in ((1, "SS, 1 issuer"), (1 + 16, "TS, 1 issuer"), (1 + 32, "SS + 4 warps spinning on an mbarrier"), (1 + 64, "SS + 8 spinning warps"),
                                (1 + 128 + 64, "SS + 20 spinning warps"), (1 + 256, "SS + a lane polling global memory"), (1 + 16 + 64, "TS + 8 spinning warps"),
                                (2, "SS, 2 issuers"), (2 + 16, "TS, 2 issuers")):
                h.test_mma_rate(M, N, n, 0, flags)
                issue, total = h.test_mma_rate(M, N, n, 0, flags)
                print(f"M={M:3d} N={N:3d} {what:45s}: issue {issue / n:6.1f} clk/mma, complete {total / n:6.1f} clk/mma per issuer", flush=True)
