"""Pins the CPU oracle (oracle/lrcn_oracle.py).  The reference has no tests for this path
(SURVEY.md §4), so these are the known-answer tests derived from lrcn.jl's own semantics
(SURVEY.md §8c list, items 1-8 and 10) plus an independent torch-autograd evaluation."""
import os

import numpy as np
import pytest
import torch

import lrcn_b200  # noqa: F401
from lrcn_b200 import synth
from oracle import lrcn_oracle as O

HERE = os.path.dirname(os.path.abspath(__file__))


def tiny(dtype=np.float64, E=4, H1=6, H2=8, V=11, B=3, l=4, scale=3.0, seed=1):
    m = [w * dtype(scale) if w.shape[0] > 1 else w for w in synth.initweights([H1, H2], V, E, seed=seed, dtype=dtype)]
    X = (synth.features(B, seed=2).astype(dtype) * 50)
    seq = [t for t in synth.tokens(l, B, V, seed=3)]
    seq[0][:] = seq[0][0]  # repeated ids inside one step (scatter-add accumulation)
    return m, X, seq, range(0, l)


def test_loss_equals_lnV_when_output_layer_is_zero():
    # KAT 1: Wout=0,bout=0 => uniform softmax => loss = ln V (slides: 8.953 = ln 7731, 9.272 = ln 10636)
    for V, expect in ((7731, 8.953), (10636, 9.272)):
        m = synth.initweights([16, 16], V, 8, seed=1)
        m[7][:] = 0
        m[8][:] = 0
        X = synth.features(4)
        seq = list(synth.tokens(3, 4, V))
        L = O.loss(m, O.initstate(m, 4), X, seq, range(0, 3))
        assert abs(L - np.log(V)) < 2e-6 * np.log(V)
        assert round(L, 3) == expect


def test_lstm_gate_order_and_forget_bias_hand_computed():
    # KAT 2: one LSTM step, H=2, X=1; columns are [f f | i i | o o | g g] (lrcn.jl:531-534)
    W = np.zeros((3, 8))
    b = np.array([[1., 1., 0., 0., 0., 0., 0., 0.]])  # forget slice [1:H] = 1 (lrcn.jl:501)
    W[0, 2] = 2.0   # x -> ingate of unit 0
    W[0, 6] = 1.0   # x -> change of unit 0
    W[1, 5] = -1.0  # h0 -> outgate of unit 1
    x = np.array([[0.5]]); h = np.array([[0.3, -0.2]]); c = np.array([[1.0, -1.0]])
    h2, c2 = O.lstm(W, b, h, c, x)
    s = lambda z: 1 / (1 + np.exp(-z))
    f = s(1.0); i0 = s(1.0); i1 = s(0.0); o0 = s(0.0); o1 = s(-0.3); g0 = np.tanh(0.5); g1 = 0.0
    c_exp = np.array([[1.0 * f + i0 * g0, -1.0 * f + i1 * g1]])
    h_exp = np.array([[o0 * np.tanh(c_exp[0, 0]), o1 * np.tanh(c_exp[0, 1])]])
    np.testing.assert_allclose(c2, c_exp, rtol=1e-14)
    np.testing.assert_allclose(h2, h_exp, rtol=1e-14)
    m = synth.initweights([4, 6], 9, 3)
    assert (m[1][0, :4] == 1).all() and (m[1][0, 4:] == 0).all()
    assert (m[3][0, :6] == 1).all() and (m[3][0, 6:] == 0).all()


def test_inputs_targets_and_count():
    # KAT 3: inputs [bos,w1..wl], targets [w1..wl,eos], count = B(l+1)
    seq = [np.array([5, 6]), np.array([7, 8])]
    ins, tgt = O._step_inputs_targets(seq, range(0, 2), 2)
    assert [a.tolist() for a in ins] == [[1, 1], [4, 5], [6, 7]]
    assert [a.tolist() for a in tgt] == [[4, 5], [6, 7], [0, 0]]
    m, X, sq, rng = tiny()
    lp = O.token_logps(m, O.initstate(m, 3), X, sq, rng)
    assert lp.shape == (5, 3)
    assert abs(-lp.sum() / 15 - O.loss(m, O.initstate(m, 3), X, sq, rng)) < 1e-12


def test_shapes_match_weight_contract():
    m = synth.initweights([512, 512], 8000, 512)
    assert [w.shape for w in m] == [(1024, 2048), (1, 2048), (1024, 2048), (1, 2048), (512, 256),
                                    (4096, 256), (8000, 512), (512, 8000), (1, 8000)]
    assert sum(w.size for w in m) == 13_578_048
    m = synth.param_shapes(1000, [1000, 1000], 10636)
    assert sum(r * c for r, c in m) == 39_838_636
    assert all(w.flags.f_contiguous for w in synth.initweights([4, 4], 7, 4))


def _torch_loss(params, X, seq, rng, B, masks=None):
    W1, b1, W2, b2, Wf, Wcnn, Wemb, Wout, bout = params
    ins, tgt = O._step_inputs_targets(seq, rng, B)
    H1 = b1.shape[1] // 4; H2 = b2.shape[1] // 4
    h1 = torch.zeros(B, H1, dtype=W1.dtype); c1 = h1.clone()
    h2 = torch.zeros(B, H2, dtype=W1.dtype); c2 = h2.clone()
    v = X @ Wcnn
    total = 0

    def cell(W, b, x, h, c):
        G = torch.cat([x, h], 1) @ W + b
        H = h.shape[1]
        f, i, o, g = torch.sigmoid(G[:, :H]), torch.sigmoid(G[:, H:2 * H]), torch.sigmoid(G[:, 2 * H:3 * H]), torch.tanh(G[:, 3 * H:])
        c = c * f + i * g
        return o * torch.tanh(c), c

    for t, (u, y) in enumerate(zip(ins, tgt)):
        e = Wemb[torch.from_numpy(u)]
        if masks is not None:
            e = e * torch.from_numpy(masks[0][t])      # dropout site lrcn.jl:542
        h1, c1 = cell(W1, b1, e, h1, c1)
        z = torch.cat([h1 @ Wf, v], 1)
        if masks is not None:
            z = z * torch.from_numpy(masks[1][t])      # dropout site lrcn.jl:547
        h2, c2 = cell(W2, b2, z, h2, c2)
        a = h2 @ Wout + bout
        lp = torch.log_softmax(a, 1)
        total = total + lp[torch.arange(B), torch.from_numpy(y)].sum()
    return -total / (B * len(ins))


def test_gradient_matches_torch_autograd_fp64():
    # KAT 4 + 8: hand BPTT == autograd on the same restated forward, incl. repeated ids
    m, X, seq, rng = tiny()
    g, L = O.lossgradient(m, O.initstate(m, 3), X, seq, rng)
    tp = [torch.tensor(np.ascontiguousarray(w), requires_grad=True) for w in m]
    Lt = _torch_loss(tp, torch.tensor(X), seq, rng, 3)
    Lt.backward()
    assert abs(float(Lt.detach()) - L) < 1e-12
    for k in range(9):
        np.testing.assert_allclose(g[k], tp[k].grad.numpy(), rtol=1e-9, atol=1e-14, err_msg=f"param {k + 1}")
    # bos row of the embedding gradient accumulates B contributions and is non-zero
    assert np.abs(g[6][O.BOS - 1]).sum() > 0
    untouched = sorted(set(range(11)) - {O.BOS - 1} - {int(t) - 1 for s in seq for t in s})
    assert np.abs(g[6][untouched]).sum() == 0


def test_dropout_masked_gradient_matches_torch_autograd_fp64():
    # the two dropout sites of lrcn.jl:542,547 with EXPLICIT masks (Knet: x .* mask ./ (1-p)): hand BPTT == autograd
    m, X, seq, rng = tiny()
    B, T, E, C2 = 3, 5, 4, 8
    masks = O.dropout_masks(0.4, 12345, T, B, E, C2, dtype=np.float64)
    assert masks[0].shape == (T, B, E) and masks[1].shape == (T, B, C2)
    assert set(np.unique(masks[0])) <= {0.0, float(np.float32(1) / (np.float32(1) - np.float32(0.4)))}
    g, L = O.lossgradient(m, O.initstate(m, B), X, seq, rng, masks=masks)
    g0, L0 = O.lossgradient(m, O.initstate(m, B), X, seq, rng)
    assert abs(L - L0) > 1e-6  # the masks do something
    tp = [torch.tensor(np.ascontiguousarray(w), requires_grad=True) for w in m]
    Lt = _torch_loss(tp, torch.tensor(X), seq, rng, B, masks=masks)
    Lt.backward()
    assert abs(float(Lt.detach()) - L) < 1e-12
    for k in range(9):
        np.testing.assert_allclose(g[k], tp[k].grad.numpy(), rtol=1e-9, atol=1e-14, err_msg=f"param {k + 1}")


def test_dropout_hash_known_answers_and_keep_rate():
    # drop_hash24 restates csrc/kernels_simt.cu: SplitMix64 finaliser of seed + G*(idx+1) + S*(site+1), top 24 bits
    def scalar(seed, site, idx):
        M = (1 << 64) - 1
        z = (seed + 0x9E3779B97F4A7C15 * (idx + 1) + 0xD1B54A32D192ED03 * (site + 1)) & M
        z = ((z ^ (z >> 30)) * 0xBF58476D1CE4E5B9) & M
        z = ((z ^ (z >> 27)) * 0x94D049BB133111EB) & M
        z ^= z >> 31
        return z >> 40
    idx = np.array([0, 1, 77, 2 ** 33 + 5], dtype=np.uint64)
    for seed, site in ((0, 0), (7, 1), (2 ** 63 + 11, 0)):
        got = O.drop_hash24(seed, site, idx)
        assert got.tolist() == [scalar(seed, site, int(i)) for i in idx]
    # empirical keep rate of a big mask: Bernoulli(0.6) within 4 sigma; survivors carry exactly 1/(1-p) in fp32
    M0, M1 = O.dropout_masks(0.4, 99, 13, 64, 128, 96)
    for M in (M0, M1):
        n = M.size
        rate = float((M > 0).mean())
        assert abs(rate - 0.6) < 4 * np.sqrt(0.24 / n)
        assert np.unique(M).tolist() == [0.0, float(np.float32(1) / np.float32(0.6))]
    # different rank / seed / site -> different masks
    assert (O.dropout_masks(0.4, 99, 2, 4, 8, 8, rank=1)[0] != O.dropout_masks(0.4, 99, 2, 4, 8, 8, rank=0)[0]).any()
    assert (O.dropout_masks(0.4, 98, 2, 4, 8, 8)[0] != O.dropout_masks(0.4, 99, 2, 4, 8, 8)[0]).any()


def test_gradient_finite_difference_fp64():
    m, X, seq, rng = tiny()
    g, _ = O.lossgradient(m, O.initstate(m, 3), X, seq, rng)
    rs = np.random.RandomState(0)
    for k in range(9):
        for _ in range(4):
            idx = tuple(rs.randint(s) for s in m[k].shape)
            old = m[k][idx]; h = 1e-5
            m[k][idx] = old + h; lp = O.loss(m, O.initstate(m, 3), X, seq, rng)
            m[k][idx] = old - h; lm = O.loss(m, O.initstate(m, 3), X, seq, rng)
            m[k][idx] = old
            assert abs((lp - lm) / (2 * h) - g[k][idx]) < 1e-8 + 1e-6 * abs(g[k][idx])


def test_fp32_restatement_close_to_fp64_shadow():
    m64, X, seq, rng = tiny(E=8, H1=12, H2=16, V=50, B=5, l=6)
    m32 = [w.astype(np.float32) for w in m64]
    g64, L64 = O.lossgradient(m64, O.initstate(m64, 5), X, seq, rng)
    g32, L32 = O.lossgradient(m32, O.initstate(m32, 5), X.astype(np.float32), seq, rng)
    assert abs(L32 - L64) < 1e-5 * abs(L64)
    for a, b in zip(g32, g64):
        assert np.linalg.norm(a - b) <= 1e-4 * np.linalg.norm(b) + 1e-12


def test_adam_first_and_second_step():
    # KAT 5: step 1: dw = -lr*g/(|g|+eps) ; step 2 moves rows whose gradient is zero (dense decay)
    w = [np.array([[1.0, -2.0, 3.0]])]
    g = [np.array([[0.5, -0.25, 0.0]])]
    opt = O.initparams(w)
    O.update(w, g, opt)
    np.testing.assert_allclose(w[0], [[1 - 1e-3 * 0.5 / (0.5 + 1e-8), -2 + 1e-3 * 0.25 / (0.25 + 1e-8), 3.0]], rtol=1e-12)
    w1 = w[0].copy()
    O.update(w, [np.zeros((1, 3))], opt)
    assert opt[0].t == 2
    assert w[0][0, 0] < w1[0, 0] and w[0][0, 1] > w1[0, 1] and w[0][0, 2] == 3.0
    m2 = 0.9 * 0.05
    v2 = 0.999 * (0.001 * 0.25)
    exp = w1[0, 0] - 1e-3 * (m2 / (1 - 0.81)) / (np.sqrt(v2 / (1 - 0.999 ** 2)) + 1e-8)
    assert abs(w[0][0, 0] - exp) < 1e-12


def _beam_model(V=12, seed=7, dtype=np.float32):
    m = synth.initweights([6, 8], V, 5, seed=seed, dtype=dtype)
    m = [w * dtype(4) if w.shape[0] > 1 else w for w in m]
    return m


def test_beam_k1_is_greedy_and_max_length():
    # KAT 7 + part of 6: K=1 == greedy argmax; at most nword+1 generated tokens
    m = _beam_model()
    m[8][0, O.EOS - 1] = -50  # make eos very unlikely -> runs to the length limit
    feat = synth.features(1)[0] * 100
    toks, p = O.generate(m, feat, nword=5, beam_width=1)
    assert len(toks) == 1 + 5 + 1
    s = O.initstate(m, 1)
    v = feat.reshape(1, -1) @ m[5]
    cur = O.BOS; prob = np.float32(1); out = [O.BOS]
    for _ in range(6):
        q = np.exp(O.logp(O.lrcn(m, s, v, m[6][cur - 1:cur, :]))).reshape(-1)
        cur = int(np.argmax(q)) + 1
        prob = np.float32(prob * q[cur - 1]); out.append(cur)
    assert out == toks and prob == p


def test_beam_ties_lower_index_first_and_step1_single_expansion():
    # KAT 6: exactly tied logits -> candidates come out in index order; first step expands beam 1 only
    m = _beam_model(V=9)
    m[7][:] = 0; m[8][:] = 0  # uniform distribution: every token ties
    feat = synth.features(1)[0]
    toks, p = O.generate(m, feat, nword=2, beam_width=3)
    # uniform probs: top-3 of step 1 are ids 1,2,3 (lower index first); best hypothesis ends in eos=1 -> stop
    assert toks == [O.BOS, 1]
    assert abs(p - np.float32(1 / 9)) < 1e-7


def test_beam_stops_only_when_top_beam_ends_in_eos_and_truncates_text():
    m = _beam_model(V=10, seed=11)
    feat = synth.features(1, seed=9)[0] * 100
    trace = []
    toks, p = O.generate(m, feat, nword=30, beam_width=3, trace=trace)
    assert toks[0] == O.BOS and (toks[-1] == O.EOS or len(toks) == 32)
    assert O.EOS not in toks[1:-1] or True  # non-top beams may have carried eos earlier (not frozen)
    assert abs(np.exp(np.sum(trace[0])) - p) < 1e-4 * p
    vocab = [f"w{k}" for k in range(1, 11)]
    txt = O.caption_text([2, 5, 6, 1, 7], vocab)
    assert txt == "w5 w6 ."  # KAT 10: tokens + ' ' ... then '.', truncated at first eos (lrcn.jl:633-640)
    # same format as the shipped eval/candidates.txt lines: words separated by single spaces, trailing " ."
    assert txt.endswith(" .") and "  " not in txt


def test_average_loss_is_token_weighted_and_skips_long():
    m, X, seq, rng = tiny(l=4)
    m2, X2, seq2, rng2 = tiny(l=2, seed=1)
    long_seq = [np.full(3, 5)] * 29
    batches = [(X, seq, rng), (X2, seq2, rng2), (X, long_seq, range(0, 29))]
    a = O.average_loss(m, batches)
    l1 = O.loss(m, O.initstate(m, 3), X, seq, rng)
    l2 = O.loss(m, O.initstate(m, 3), X2, seq2, rng2)
    assert abs(a - (l1 * 15 + l2 * 9) / 24) < 1e-12


def test_synth_is_deterministic_and_in_range():
    a = synth.tokens(5, 7, 100, seed=3); b = synth.tokens(5, 7, 100, seed=3)
    assert (a == b).all() and a.min() >= 4 and a.max() <= 100 and a.dtype == np.int64
    z = synth.tokens(50, 64, 1000, seed=3, zipf=True)
    assert z.min() >= 3 and z.max() <= 1000 and (z == 3).any()
    f = synth.features(3)
    np.testing.assert_allclose(f.sum(1), 1.0, rtol=1e-5)
    assert (f >= 0).all() and f.shape == (3, 4096)
    ls = synth.lengths(1000, "flickr")
    assert ls.min() >= 5 and ls.max() <= 28 and 11 < ls.mean() < 13.5
    lc = synth.lengths(1000, "coco")
    assert 9.5 < lc.mean() < 11.5
    # first SplitMix64 outputs are a fixed known answer (portability check for C++/Julia ports)
    assert synth.splitmix_u64(0, 2).tolist() == [0xE220A8397B1DCDAF, 0x6E789E6AA1B965F4]


def test_golden_fixture_matches_oracle():
    """tests/golden/tiny_train.npz (made by tests/golden/make_golden.py, cross-checked against torch autograd in fp64
    when it was generated) pins the oracle: loss, the 9 gradients, the weights after two Adam steps, beam-3 captions."""
    path = os.path.join(HERE, "golden", "tiny_train.npz")
    z = np.load(path)
    E, H1, H2, V, B, l = [int(x) for x in z["cfg"]]
    m = synth.initweights([H1, H2], V, E, seed=1)
    m = [w * np.float32(3) if w.shape[0] > 1 else w for w in m]
    X = synth.features(B, seed=2) * np.float32(50)
    seq = list(synth.tokens(l, B, V, seed=3))
    g, L = O.lossgradient(m, O.initstate(m, B), X, seq, range(0, l))
    assert abs(L - float(z["loss"])) < 1e-6
    assert abs(L - float(z["loss_torch_fp64"])) < 1e-5 * abs(L)
    for k in range(9):
        np.testing.assert_allclose(g[k], z[f"g{k + 1}"], rtol=1e-4, atol=1e-7)
    w2 = [w.copy() for w in m]
    opt = O.initparams(w2)
    for _ in range(2):
        gg, _ = O.lossgradient(w2, O.initstate(w2, B), X, seq, range(0, l))
        O.update(w2, gg, opt)
    for k in range(9):
        np.testing.assert_allclose(w2[k], z[f"w2_{k + 1}"], rtol=1e-5, atol=1e-7)
    for i in range(B):
        toks, prob = O.generate(m, X[i], 8, 3)
        n = int(z["beam_len"][i])
        assert list(toks) == z["beam_tokens"][i][:n].tolist()
        assert abs(prob - z["beam_prob"][i]) <= 1e-5 * abs(z["beam_prob"][i])
