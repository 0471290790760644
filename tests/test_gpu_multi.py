"""Multi-GPU parity on real GPUs (skipped below 2 visible GPUs; run with `gpurun --gpus N`).

(1) single-process group (cfg.n_gpus > 1, SURVEY 8b "the library drives all 1/2/4/8 GPUs itself"): ONE handle, global
    batches; the sharded gradient == the oracle's gradient of the global batch, the sharded-Adam weights == the oracle's
    Adam step, beam search sharded by image == the single-GPU result;
(2) one process per GPU (torchrun + CUDA IPC, the way bench.py is launched): tools/dp_check.py as a subprocess -- gradient ==
    oracle gradient of the global batch, Adam state == oracle's, bit-identical replicas after 4 steps.
Both write their report to gpurun_out/ (copied to profiles/ by the builder)."""
import json
import os
import subprocess
import sys

import numpy as np
import pytest
import torch

import lrcn_b200  # noqa: F401
from lrcn_b200 import abi, synth
from oracle import lrcn_oracle as O

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
NGPU = torch.cuda.device_count() if torch.cuda.is_available() else 0
needs_multi = pytest.mark.skipif(NGPU < 2, reason="needs >= 2 GPUs on one node")


def relerr(a, b):
    return float(np.linalg.norm(np.asarray(a, np.float64) - np.asarray(b, np.float64)) / (np.linalg.norm(np.asarray(b, np.float64)) + 1e-30))


def report(name, obj):
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    with open(os.path.join(ROOT, "gpurun_out", name), "a") as f:
        f.write(json.dumps(obj) + "\n")


def group_sizes():
    return [n for n in (2, 4, 8) if n <= NGPU]


@needs_multi
@pytest.mark.parametrize("prec", [abi.PREC_FP32, abi.PREC_BF16X3])
def test_single_process_group_matches_oracle_and_single_gpu(prec):
    E, H1, H2, V, l = 64, 64, 64, 300, 5
    model = synth.initweights([H1, H2], V, E, seed=1)
    model = [w * np.float32(3) if w.shape[0] > 1 else w for w in model]
    feats = synth.features(32, seed=2) * np.float32(50)
    ids = np.arange(1, 33, dtype=np.int64)
    for N in group_sizes():
        Bg = 6 * N + 1  # ragged split: shard sizes differ by one
        img = synth.image_ids(Bg, 32)
        tok = synth.tokens(l, Bg, V)
        X = feats[img - 1]
        g_ref, L_ref = O.lossgradient(model, O.initstate(model, Bg), X, list(tok), range(0, l))
        lp_ref = O.token_logps(model, O.initstate(model, Bg), X, list(tok), range(0, l))
        cfg = abi.default_config(embed=E, hidden1=H1, hidden2=H2, vocab=V, max_batch=Bg, max_len=8, max_gen_rows=12 * N, precision=prec, n_gpus=N)
        for i in range(N):
            cfg.device_ids[i] = i
        with abi.Handle(cfg) as h:
            h.set_model(model)
            h.load_features(0, ids, feats)
            h.load_features(1, ids, feats)
            s, n = h.loss(0, img, tok)
            assert n == Bg * (l + 1) and abs(-s / n - L_ref) < 1e-4 * abs(L_ref)
            np.testing.assert_allclose(h.token_logps(l, Bg), lp_ref, rtol=1e-4, atol=1e-5)
            L = h.grad(0, img, tok)
            assert abs(L - L_ref) < 1e-4 * abs(L_ref)
            errs = [relerr(h.get_grad(k), g_ref[k - 1]) for k in range(1, 10)]
            assert max(errs) < 1e-4, errs
            # dropout: every shard draws its own masks (the rank is mixed into the seed; rows are local to the shard)
            masks = [np.zeros((l + 1, Bg, E), np.float32), np.zeros((l + 1, Bg, H2), np.float32)]
            off = 0
            for r in range(N):
                b = Bg // N + (1 if r < Bg % N else 0)
                m0, m1 = O.dropout_masks(0.4, 99, l + 1, b, E, H2, rank=r)
                masks[0][:, off:off + b], masks[1][:, off:off + b] = m0, m1
                off += b
            gd_ref, Ld_ref = O.lossgradient(model, O.initstate(model, Bg), X, list(tok), range(0, l), masks=masks)
            Ld = h.grad(0, img, tok, 0.4, 99)
            assert abs(Ld - Ld_ref) < 1e-4 * abs(Ld_ref)
            errs_d = [relerr(h.get_grad(k), gd_ref[k - 1]) for k in range(1, 10)]
            assert max(errs_d) < 1e-4, errs_d
            # training: sharded Adam over peer memory == the oracle's Adam on the global-batch gradient
            ref = [w.copy() for w in model]
            opt = O.initparams(ref)
            for step in range(3):
                L_o = O.train_step(ref, opt, X, list(tok), range(0, l))
                L_g = h.train_step(0, img, tok)
                assert abs(L_g - L_o) < 1e-4 * abs(L_o), (N, step)
            assert h.get_adam_step() == 3
            derr = [relerr(h.get_param(k) - model[k - 1], ref[k - 1] - model[k - 1]) for k in range(1, 10)]
            assert max(derr) < 5e-3, derr
            merr = max(relerr(h.get_adam_state(k, 0), opt[k - 1].fstm) for k in range(1, 10))
            verr = max(relerr(h.get_adam_state(k, 1), opt[k - 1].scndm) for k in range(1, 10))
            assert merr < 2e-3 and verr < 2e-3
            # the summed gradient of the last train step is gathered from its owners
            assert all(np.isfinite(h.get_grad(k)).all() for k in range(1, 10))
            # the epoch call on the group (each member stages its columns of the global batch on its own GPU) == per-step calls
            h.set_model(model)
            h.set_adam_step(0)
            for k in range(1, 10):
                h.set_adam_state(k, 0, np.zeros_like(model[k - 1])); h.set_adam_state(k, 1, np.zeros_like(model[k - 1]))
            blens = [l, 2, 4]
            seq_e = np.concatenate([synth.tokens(bl, Bg, V, seed=40 + i) for i, bl in enumerate(blens)])
            img_e = np.stack([synth.image_ids(Bg, 32, seed=50 + i) for i in range(3)])
            ep = h.train_epoch(0, seq_e, img_e, blens, np.array([2, 0, 1]), 0.4, seed=5)
            w_ep = [h.get_param(k) for k in range(1, 10)]
            h.set_model(model)
            h.set_adam_step(0)
            for k in range(1, 10):
                h.set_adam_state(k, 0, np.zeros_like(model[k - 1])); h.set_adam_state(k, 1, np.zeros_like(model[k - 1]))
            st = np.concatenate([[0], np.cumsum(blens)])
            per = [h.train_step(0, img_e[b], seq_e[st[b]:st[b] + blens[b]], 0.4, 5 + i) for i, b in enumerate([2, 0, 1])]
            np.testing.assert_allclose(ep, per, rtol=1e-6)
            assert max(relerr(a - model[k], h.get_param(k + 1) - model[k]) for k, a in enumerate(w_ep)) < 1e-3
            s_e, n_e = h.loss_epoch(0, seq_e, img_e, blens)
            per_l = [h.loss(0, img_e[b], seq_e[st[b]:st[b] + blens[b]]) for b in range(3)]
            assert n_e == sum(p_[1] for p_ in per_l) and abs(s_e - sum(p_[0] for p_ in per_l)) < 1e-5 * abs(s_e)
            # generation: images sharded over the GPUs, no collective; same captions as one GPU
            h.set_model(model)
            toks, lens, prob, _ = h.beam_search(1, ids[:4 * N + 1], 3, 6)
        with abi.Handle(abi.default_config(embed=E, hidden1=H1, hidden2=H2, vocab=V, max_batch=8, max_len=8, max_gen_rows=12, precision=prec)) as h1:
            h1.set_model(model)
            h1.load_features(1, ids, feats)
            t1, l1, p1, _ = h1.beam_search(1, ids[:4 * N + 1], 3, 6)
        assert np.array_equal(lens, l1) and np.array_equal(toks, t1) and np.allclose(prob, p1, rtol=1e-5)
        report("dp_group_check.jsonl", dict(mode="single-process group", n_gpus=N, precision=int(prec), global_batch=Bg, loss=L, loss_oracle=L_ref,
                                            max_grad_relerr=max(errs), max_grad_relerr_pdrop04=max(errs_d), max_update_relerr=max(derr),
                                            adam_m_relerr=merr, adam_v_relerr=verr, beam_identical=True))


@needs_multi
def test_single_process_group_full_size_c2():
    """configs[1] on the single-process group: 256 captions per GPU, E = H = 512, V = 7731; the global gradient against the
    oracle run in 64-row slices."""
    N = group_sizes()[0]
    E = H = 512
    V, b, l = 7731, 64, 9
    Bg = b * N
    model = synth.initweights([H, H], V, E, seed=1)
    model = [w * np.float32(1.5) if w.shape[0] > 1 else w for w in model]
    feats = synth.features(128, seed=2) * np.float32(50)
    ids = np.arange(1, 129, dtype=np.int64)
    img = synth.image_ids(Bg, 128)
    tok = synth.tokens(l, Bg, V, zipf=True)
    g_sum, L_sum = None, 0.0
    for s_ in range(N):
        sl = slice(b * s_, b * (s_ + 1))
        g, L = O.lossgradient(model, O.initstate(model, b), feats[img[sl] - 1], list(tok[:, sl]), range(0, l))
        g_sum = g if g_sum is None else [a + c for a, c in zip(g_sum, g)]
        L_sum += L
    g_ref = [a / np.float32(N) for a in g_sum]
    cfg = abi.default_config(embed=E, hidden1=H, hidden2=H, vocab=V, max_batch=Bg, max_len=28, max_gen_rows=8, precision=abi.PREC_BF16X3, n_gpus=N)
    with abi.Handle(cfg) as h:
        h.set_model(model)
        h.load_features(0, ids, feats)
        L = h.grad(0, img, tok)
        assert abs(L - L_sum / N) < 1e-4 * abs(L)
        errs = [relerr(h.get_grad(k), g_ref[k - 1]) for k in range(1, 10)]
        assert max(errs) < 1e-4, errs
    report("dp_group_check.jsonl", dict(mode="single-process group, configs[1] dims", n_gpus=N, global_batch=Bg, max_grad_relerr=max(errs)))


@needs_multi
@pytest.mark.parametrize("N", [2, 4, 8])
def test_one_process_per_gpu_dp_check(N):
    """tools/dp_check.py under torchrun: the launch mode of bench.py --gpus N (CUDA IPC peer memory)."""
    if N > NGPU:
        pytest.skip(f"needs {N} GPUs")
    port = 29500 + N
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={N}", "--master-addr", "127.0.0.1", "--master-port", str(port),
           os.path.join(ROOT, "tools", "dp_check.py")]
    out = subprocess.run(cmd, capture_output=True, text=True, timeout=900, cwd=ROOT)
    lines = [x for x in out.stdout.splitlines() if x.startswith("dp_check")]
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    with open(os.path.join(ROOT, "gpurun_out", f"dp_check_{N}gpu.log"), "w") as f:
        f.write("\n".join(lines) + "\n")
    assert out.returncode == 0, out.stdout[-3000:] + out.stderr[-3000:]
    assert len(lines) == 2 and all(x.endswith("OK") for x in lines), lines
