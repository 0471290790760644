"""generation throughput versus images in flight and beam width (E=H=512, V=10000, nword=30,
untrained weights => every caption runs the full 31 steps; and the eos-timed COCO-shaped model beside it).  Prints a markdown table for profiles/."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import lrcn_b200  # noqa: E402,F401
from lrcn_b200 import abi  # noqa: E402
import bench  # noqa: E402

if __name__ == "__main__":
    print("| images in flight | beam K | rows | worst case (31 steps): ms / batch | captions/s | beam-row steps/s | us / step | COCO-shaped: ms / batch | captions/s | mean steps |")
    print("|---:|---:|---:|---:|---:|---:|---:|---:|---:|---:|")
    for K in (1, 3, 5, 10):
        for n in (32, 128, 512, 1024, 4096):
            if n * K > 16384:
                continue
            r = bench.beam_leg(0, 0, 1, abi.PREC_BF16X3, lambda x: x, lambda: None, False, n_img=n, K=K)
            c = bench.beam_leg(0, 0, 1, abi.PREC_BF16X3, lambda x: x, lambda: None, True, n_img=n, K=K)
            print(f"| {n} | {K} | {n * K} | {r['ms_per_batch']:.2f} | {r['value']:.0f} | {r['row_steps_per_s']:.3g} | {1e3 * r['ms_per_batch'] / 31:.0f} | "
                  f"{c['ms_per_batch']:.2f} | {c['value']:.0f} | {c['mean_decode_steps']:.1f} |", flush=True)
