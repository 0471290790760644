"""GPU probe: where does the bf16x3 pair GEMM spend its time?  Times each shape with the diagnostic switches of
lrcn_test_gemm_time (1 = no epilogue stores, 2 = no TMA loads after the first ring fill, 4 = no MMAs).
Usage: python tools/probe_gemm_time.py"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import lrcn_b200  # noqa: F401,E402
from lrcn_b200 import abi  # noqa: E402

SHAPES = [
    ("vocab fwd", 1, 1, 3328, 7731, 512, 0),
    ("vocab fwd K=1024", 1, 1, 3328, 7731, 1024, 0),
    ("vocab fwd K=2048", 1, 1, 3328, 7731, 2048, 0),
    ("x-gates", 1, 1, 3328, 2048, 512, 0),
    ("dWout", 0, 0, 7731, 512, 3328, 0),
    ("dh2", 1, 0, 3328, 512, 7731, 0),
    ("dZ", 1, 0, 3328, 512, 2048, 0),
    ("dW2x", 0, 0, 2048, 512, 3328, 0),
    ("Z", 1, 1, 3328, 256, 512, 0),
    ("v", 1, 1, 256, 256, 4096, 0),
    ("dWcnn", 0, 0, 256, 4096, 256, 0),
    ("square 8192", 1, 1, 8192, 8192, 8192, 0),
]

if __name__ == "__main__":
    only = sys.argv[sys.argv.index("--only") + 1] if "--only" in sys.argv else None
    dbgs = [int(sys.argv[sys.argv.index("--dbg") + 1])] if "--dbg" in sys.argv else [0, 1, 2, 4, 3, 6]
    cfg = abi.default_config(embed=64, hidden1=64, hidden2=64, vocab=100, max_batch=4, max_len=2, max_gen_rows=4, precision=1)
    with abi.Handle(cfg, hooks=True) as h:
        for (name, aK, bK, M, N, K, sh) in SHAPES:
            if (only is None and name.startswith("square")) or (only is not None and not name.startswith(only)):
                continue
            line = f"{name:18s} {M}x{N}x{K} aK={aK} bK={bK}:"
            for dbg in dbgs:
                ms = h.test_gemm_time(aK, bK, M, N, K, bool(sh), 5 if M * N * K > 1e11 else 20, dbg)
                tf = 2.0 * M * N * K / ms / 1e9
                line += f"  dbg{dbg}={ms*1e3:7.1f}us ({tf:5.0f}TF)"
            print(line, flush=True)
