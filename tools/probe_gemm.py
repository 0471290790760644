"""GPU probe: run every GEMM variant (fp32 CUDA-core and tcgen05 bf16x3; all operand major combinations)
against an fp64 numpy product and print relative errors.  Each case runs in its own subprocess so a
faulting kernel cannot poison the others.  Usage: python tools/probe_gemm.py [--one prec aK bK M N K]"""
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def one(prec, aK, bK, M, N, K, beta, bias):
    import numpy as np
    import lrcn_b200  # noqa: F401
    from lrcn_b200 import abi
    rs = np.random.RandomState(M * 7 + N * 3 + K)
    A = rs.standard_normal((M, K)).astype(np.float32)
    B = rs.standard_normal((K, N)).astype(np.float32)
    bvec = rs.standard_normal(N).astype(np.float32) if bias else None
    C0 = rs.standard_normal((M, N)).astype(np.float32) if beta else None
    ref = A.astype(np.float64) @ B.astype(np.float64)
    if bias:
        ref += bvec
    if beta:
        ref += C0
    cfg = abi.default_config(embed=64, hidden1=64, hidden2=64, vocab=100, max_batch=4, max_len=2, max_gen_rows=4, precision=prec)
    with abi.Handle(cfg, hooks=True) as h:
        out = h.test_gemm(prec, aK, bK, A if aK else np.ascontiguousarray(A.T), np.ascontiguousarray(B.T) if bK else B, bvec, C0)
    err = np.linalg.norm(out - ref) / np.linalg.norm(ref)
    mx = np.abs(out - ref).max()
    print(f"prec={prec} aK={aK} bK={bK} M={M} N={N} K={K} beta={beta} bias={bias} rel={err:.3e} maxabs={mx:.3e}", flush=True)


if __name__ == "__main__":
    if len(sys.argv) > 1 and sys.argv[1] == "--one":
        a = [int(x) for x in sys.argv[2:10]]
        one(*a)
        sys.exit(0)
    shapes = [(128, 128, 64), (128, 128, 256), (256, 384, 512), (200, 136, 520), (64, 2048, 512), (1344, 8000, 512), (1000, 500, 7731)]
    envs = [{"LRCN_GEMM_2CTA": "0", "LRCN_GEMM_BN": "128"}, {"LRCN_GEMM_2CTA": "0", "LRCN_GEMM_BN": "256"}, {"LRCN_GEMM_2CTA": "1"}]
    if "--2cta" in sys.argv:
        envs = [{"LRCN_GEMM_2CTA": "1"}]
    if "--quick" in sys.argv:
        shapes = [(200, 136, 520), (1344, 8000, 512), (1000, 500, 7731), (3328, 2048, 512), (257, 300, 64), (1536, 520, 4096), (3328, 512, 7731)]
    for env in envs:
        print("== env", env, flush=True)
        for prec in (0, 1):
            for (M, N, K) in shapes:
                for aK in (1, 0):
                    for bK in (1, 0):
                        if prec == 0 and env.get("LRCN_GEMM_BN") != "128":
                            continue
                        e = dict(os.environ)
                        e.update(env)
                        try:
                            r = subprocess.run([sys.executable, __file__, "--one", str(prec), str(aK), str(bK), str(M), str(N), str(K),
                                                str(int(M % 3 == 0)), str(int(N % 5 == 0))], env=e, capture_output=True, text=True, timeout=120)
                            sys.stdout.write(r.stdout)
                            if r.returncode != 0:
                                print(f"FAIL prec={prec} aK={aK} bK={bK} {M}x{N}x{K}: rc={r.returncode} {r.stderr[-400:]}", flush=True)
                        except subprocess.TimeoutExpired:
                            print(f"TIMEOUT prec={prec} aK={aK} bK={bK} {M}x{N}x{K}", flush=True)
