"""Generates tests/golden/tiny_train.npz.

The reference (Julia 0.5 + Knet, un-pinned, not installable here) cannot be executed, so these vectors
are NOT outputs of the reference: they are outputs of oracle/lrcn_oracle.py (the NumPy restatement of lrcn.jl:489-581,
585-678) on seeded synthetic inputs, cross-checked at generation time against an independent torch-autograd evaluation
of the same forward.  They pin the oracle against drift and give the GPU path a fixed target that does not execute
oracle/ code.  Re-run from the repo root:  python tests/golden/make_golden.py
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
import lrcn_b200  # noqa: E402,F401
from lrcn_b200 import synth  # noqa: E402
from oracle import lrcn_oracle as O  # noqa: E402

E, H1, H2, V, B, l = 24, 16, 32, 57, 6, 5


def inputs():
    m = synth.initweights([H1, H2], V, E, seed=1)
    m = [w * np.float32(3) if w.shape[0] > 1 else w for w in m]
    X = synth.features(B, seed=2) * np.float32(50)
    seq = list(synth.tokens(l, B, V, seed=3))
    return m, X, seq


def torch_loss(m, X, seq):
    import torch
    t = [torch.tensor(w, dtype=torch.float64, requires_grad=True) for w in m]
    x = torch.tensor(X, dtype=torch.float64)

    def lstm(W, b, h, c, inp):
        g = torch.cat([inp, h], 1) @ W + b
        H = h.shape[1]
        f, i, o, ch = torch.sigmoid(g[:, :H]), torch.sigmoid(g[:, H:2 * H]), torch.sigmoid(g[:, 2 * H:3 * H]), torch.tanh(g[:, 3 * H:])
        c2 = c * f + i * ch
        return o * torch.tanh(c2), c2

    h1 = c1 = torch.zeros(B, H1, dtype=torch.float64)
    h2 = c2 = torch.zeros(B, H2, dtype=torch.float64)
    v = x @ t[5]
    toks = [np.full(B, 2)] + [np.asarray(s) for s in seq]
    tgts = [np.asarray(s) for s in seq] + [np.full(B, 1)]
    total = 0.0
    for ti, yo in zip(toks, tgts):
        e = t[6][torch.tensor(ti - 1)]
        h1, c1 = lstm(t[0], t[1], h1, c1, e)
        z = torch.cat([h1 @ t[4], v], 1)
        h2, c2 = lstm(t[2], t[3], h2, c2, z)
        lp = torch.log_softmax(h2 @ t[7] + t[8], 1)
        total = total + lp[torch.arange(B), torch.tensor(yo - 1)].sum()
    loss = -total / (B * (l + 1))
    loss.backward()
    return float(loss), [w.grad.numpy() for w in t]


if __name__ == "__main__":
    m, X, seq = inputs()
    g, L = O.lossgradient(m, O.initstate(m, B), X, seq, range(0, l))
    Lt, gt = torch_loss(m, X, seq)
    assert abs(L - Lt) < 1e-5 * abs(Lt), (L, Lt)
    for k in range(9):
        assert np.linalg.norm(g[k] - gt[k]) < 2e-4 * np.linalg.norm(gt[k]) + 1e-9, k
    # two Adam steps
    w2 = [w.copy() for w in m]
    opt = O.initparams(w2)
    for _ in range(2):
        gg, _ = O.lossgradient(w2, O.initstate(w2, B), X, seq, range(0, l))
        O.update(w2, gg, opt)
    # beam search, K = 3, nword = 8, all 6 images (untrained weights: no hypothesis ends early, 10 tokens each)
    beams = [O.generate(m, X[i], 8, 3) for i in range(B)]
    out = {"cfg": np.array([E, H1, H2, V, B, l]), "loss": np.float64(L), "loss_torch_fp64": np.float64(Lt)}
    for k in range(9):
        out[f"g{k + 1}"] = g[k]
        out[f"w2_{k + 1}"] = w2[k]
    out["beam_tokens"] = np.array([np.pad(np.asarray(b[0], dtype=np.int64), (0, 10 - len(b[0]))) for b in beams])
    out["beam_len"] = np.array([len(b[0]) for b in beams])
    out["beam_prob"] = np.array([b[1] for b in beams], dtype=np.float32)
    np.savez_compressed(os.path.join(HERE, "tiny_train.npz"), **out)
    print("wrote tiny_train.npz: loss", L, "torch fp64", Lt, "beam lens", out["beam_len"])
