"""GPU parity at the sizes BASELINE.json's configs name, against the CPU oracle.

The small-shape parity tests (test_gpu_parity.py) cannot reach the code paths that only exist at full size: the 2-CTA
whole-tile / stream-K mix of the vocab projection at V = 7731 / 8000 / 10636, softmax_ce_fused<4> / <6> on a padded ldV, the
persistent LSTM kernels with 8 resident k-blocks and 4 m-tiles, the per-step H = 1000 kernels of coco_2f with C = 500, and
the top-K threshold select at V = 10 000.  One oracle lossgradient at these sizes costs 1-3 s of numpy/OpenBLAS.
Bars (north_star): loss and the nine gradients within 1e-4 relative; >= 99 % identical captions, log-probs within 1e-4;
integer selection bit-exact on identical inputs."""
import os

import numpy as np
import pytest

import lrcn_b200  # noqa: F401
from lrcn_b200 import abi, synth
from oracle import lrcn_oracle as O

pytestmark = pytest.mark.gpu

PRECS = [int(x) for x in os.environ.get("LRCN_TEST_PRECS", "0,1").split(",")]
RTOL = 1e-4


def relerr(a, b):
    return float(np.linalg.norm(np.asarray(a, np.float64) - np.asarray(b, np.float64)) / (np.linalg.norm(np.asarray(b, np.float64)) + 1e-30))


def make(E, H, V, B, l, n_img=128, scale=1.5, zipf=True, seed=1):
    model = synth.initweights([H, H], V, E, seed=seed)
    model = [w * np.float32(scale) if w.shape[0] > 1 else w for w in model]  # livelier gates than xavier alone
    feats = synth.features(n_img, seed=2) * np.float32(50)
    ids = np.arange(1, n_img + 1, dtype=np.int64)
    img = synth.image_ids(B, n_img, seed=5)
    tok = synth.tokens(l, B, V, seed=3, zipf=zipf)
    return model, feats, ids, img, tok


def handle(E, H, V, B, l, prec, **kw):
    return abi.Handle(abi.default_config(embed=E, hidden1=H, hidden2=H, vocab=V, max_batch=B, max_len=max(l, 1), max_gen_rows=kw.pop("gen_rows", 8),
                                         precision=prec, use_graphs=kw.pop("graphs", 1)))


def check_grads(h, g_ref, what):
    for k in range(1, 10):
        assert relerr(h.get_grad(k), g_ref[k - 1]) < RTOL, f"{what}: gradient of param {k}"


@pytest.mark.parametrize("prec", PRECS)
def test_c1_flickr8k_shape_full_size(prec):
    """BASELINE configs[0] / SURVEY C1: fc7 4096-d, E = H = 512, V = 8000, B = 64, l = 20 (1344 tokens per step)."""
    E = H = 512
    V, B, l = 8000, 64, 20
    model, feats, ids, img, tok = make(E, H, V, B, l)
    X = feats[img - 1]
    g_ref, L_ref = O.lossgradient(model, O.initstate(model, B), X, list(tok), range(0, l))
    lp_ref = O.token_logps(model, O.initstate(model, B), X, list(tok), range(0, l))
    with handle(E, H, V, B, l, prec) as h:
        h.set_model(model)
        h.load_features(0, ids, feats)
        s, n = h.loss(0, img, tok)
        assert n == B * (l + 1) and abs(-s / n - L_ref) < RTOL * abs(L_ref)
        np.testing.assert_allclose(h.token_logps(l, B), lp_ref, rtol=RTOL, atol=2e-5)
        L = h.grad(0, img, tok)
        assert abs(L - L_ref) < RTOL * abs(L_ref)
        check_grads(h, g_ref, "C1")
        # dropout at the reference's training setting on the full-size kernels
        masks = O.dropout_masks(0.4, 777, l + 1, B, E, H)
        gd_ref, Ld_ref = O.lossgradient(model, O.initstate(model, B), X, list(tok), range(0, l), masks=masks)
        Ld = h.grad(0, img, tok, 0.4, 777)
        assert abs(Ld - Ld_ref) < RTOL * abs(Ld_ref)
        check_grads(h, gd_ref, "C1 pdrop=0.4")


def test_c2_flickr30k_shape_b256_in_four_oracle_slices():
    """BASELINE configs[1] / SURVEY C2, the bench workload: E = H = 512, V = 7731, 256 captions, l = 12.  The oracle runs the
    batch as four 64-row slices: the gradient of a mean over 256 captions is the mean of the four slice gradients.  Also the
    64-row slice alone (one m-tile of the persistent LSTM kernels, other tile counts in every GEMM)."""
    if abi.PREC_BF16X3 not in PRECS:
        pytest.skip("tcgen05 path only")
    E = H = 512
    V, B, l = 7731, 256, 12
    model, feats, ids, img, tok = make(E, H, V, B, l, n_img=512)
    g_sum, L_sum, lps = None, 0.0, []
    slices = []
    for s_ in range(4):
        sl = slice(64 * s_, 64 * (s_ + 1))
        Xs, ts = feats[img[sl] - 1], list(tok[:, sl])
        g, L = O.lossgradient(model, O.initstate(model, 64), Xs, ts, range(0, l))
        slices.append((g, L))
        g_sum = g if g_sum is None else [a + b for a, b in zip(g_sum, g)]
        L_sum += L
    g_ref = [a / np.float32(4) for a in g_sum]
    L_ref = L_sum / 4
    with handle(E, H, V, B, 28, abi.PREC_BF16X3) as h:
        h.set_model(model)
        h.load_features(0, ids, feats)
        L = h.grad(0, img, tok)
        assert abs(L - L_ref) < RTOL * abs(L_ref)
        check_grads(h, g_ref, "C2 B=256")
        s, n = h.loss(0, img, tok)
        assert abs(-s / n - L_ref) < RTOL * abs(L_ref)
        g64, L64 = slices[0]
        L = h.grad(0, img[:64], np.ascontiguousarray(tok[:, :64]))
        assert abs(L - L64) < RTOL * abs(L64)
        check_grads(h, g64, "C2 64-row slice")
        # one Adam step on the 256-caption gradient (Knet Adam, lrcn.jl:394,402): compare the UPDATE with the oracle's
        ref = [w.copy() for w in model]
        O.update(ref, g_ref, O.initparams(ref))
        h.train_step(0, img, tok)
        for k in (1, 3, 5, 8, 9):
            d_ref = ref[k - 1] - model[k - 1]
            assert relerr(h.get_param(k) - model[k - 1], d_ref) < 5e-3, f"Adam update of param {k}"


@pytest.mark.parametrize("prec", PRECS)
def test_c4_coco_2f_reference_default_dims(prec):
    """BASELINE configs[3] / SURVEY C4: the reference's own defaults --hidden 1000 1000 --embed 1000 (lrcn.jl:39-40),
    C = 500, V = 10636 (padded ldV), COCO-shaped l = 10, B = 32."""
    E = H = 1000
    V, B, l = 10636, 32, 10
    model, feats, ids, img, tok = make(E, H, V, B, l, n_img=64)
    X = feats[img - 1]
    g_ref, L_ref = O.lossgradient(model, O.initstate(model, B), X, list(tok), range(0, l))
    with handle(E, H, V, B, l, prec) as h:
        h.set_model(model)
        h.load_features(0, ids, feats)
        L = h.grad(0, img, tok)
        assert abs(L - L_ref) < RTOL * abs(L_ref)
        check_grads(h, g_ref, "C4")
        masks = O.dropout_masks(0.4, 5, l + 1, B, E, H)
        gd_ref, Ld_ref = O.lossgradient(model, O.initstate(model, B), X, list(tok), range(0, l), masks=masks)
        Ld = h.grad(0, img, tok, 0.4, 5)
        assert abs(Ld - Ld_ref) < RTOL * abs(Ld_ref)
        check_grads(h, gd_ref, "C4 pdrop=0.4")


@pytest.mark.parametrize("K", [3, 5])
def test_c3_coco_shaped_beam_search_v10k(K):
    """BASELINE configs[2] / SURVEY C3: E = H = 512, V = 10000, beam 3 and 5, nword 30, eos-biased output bias so decodes end
    after a COCO-like number of steps (synth.eos_timed_model: mean ~11 generated tokens at K = 3).  >= 99 % identical captions, path probability and token log-probs within 1e-4
    (lrcn.jl:644-678 semantics: fp32 probability products, ties to the lower index, finished beams not frozen)."""
    if abi.PREC_BF16X3 not in PRECS:
        pytest.skip("tcgen05 path only")
    E = H = 512
    V, n_img, nword = 10000, 100, 30
    model = synth.eos_timed_model([H, H], V, E)  # eos overtakes the other words after an image-dependent number of steps
    feats = synth.features(n_img, seed=9) * np.float32(100)
    ids = np.arange(1, n_img + 1, dtype=np.int64)
    with handle(E, H, V, 4, 3, abi.PREC_BF16X3, gen_rows=n_img * K) as h:
        h.set_model(model)
        h.load_features(1, ids, feats)
        toks, lens, prob, lps = h.beam_search(1, ids, K, nword)
    same = 0
    for i in range(n_img):
        trace = []
        ref_t, ref_p = O.generate(model, feats[i], nword, K, trace=trace)
        got = toks[i, :lens[i]].tolist()
        if got == ref_t:
            same += 1
            assert abs(prob[i] - ref_p) <= 2e-4 * ref_p
            np.testing.assert_allclose(lps[i, :lens[i] - 1], np.array(trace[0], np.float32), rtol=1e-4, atol=1e-4)
    assert lens.max() > lens.min() and 4 <= lens.mean() <= 28, f"decode lengths {lens.min()}..{lens.max()} do not exercise the recurrence"
    assert same >= int(np.ceil(0.99 * n_img)), f"only {same}/{n_img} captions identical at V=10000, K={K}"


@pytest.mark.parametrize("R", [48, 2304])
@pytest.mark.parametrize("K", [1, 3, 5, 10])
def test_production_topk_kernel_from_logits_v10k(K, R):
    """The kernels generation runs -- beam_row_topk3_kernel (a CTA per row, R = 48) and beam_row_topk4_kernel (a warp per row,
    chosen from 2048 rows up: R = 2304) -- from LOGITS at V = 10 000 (lrcn.jl:652-661): near-uniform rows
    (threshold select with many near-ties), rows with exact ties (lower index wins, lrcn.jl:655), an all-equal row (more
    candidates than the list holds: exact fallback), peaked rows.  Tokens bit-exact wherever the top-(K+1) probabilities are
    distinct in fp32; scores = prob * parent in fp32 within 1e-6 relative."""
    rs = np.random.RandomState(K)
    V = 10000
    logits = (rs.standard_normal((R, V)) * 0.02).astype(np.float32)        # untrained-model-like: near uniform
    logits[8:16] = (rs.standard_normal((8, V)) * 3).astype(np.float32)      # peaked
    logits[16, 123] = logits[16, 4567] = logits[16, 9999] = 5.0              # exact three-way tie at the top
    logits[17, :] = 0.25                                                     # all equal: fallback path, indices 0..K-1
    logits[18, 7] = logits[18, 8] = -0.5                                     # tie far from the top: must not disturb
    logits[19, V - 1] = 9.0                                                  # winner in the scalar tail
    parent = rs.uniform(0.05, 1.0, R).astype(np.float32)
    with abi.Handle(abi.default_config(embed=64, hidden1=64, hidden2=64, vocab=V, max_batch=4, max_len=2, max_gen_rows=8, precision=abi.PREC_FP32),
                    hooks=True) as h:
        tok, sc, lp = h.test_beam_topk_logits(logits, parent, K)
    l64 = logits.astype(np.float64)
    lp_ref = l64 - l64.max(1, keepdims=True)
    lp_ref = lp_ref - np.log(np.exp(lp_ref).sum(1, keepdims=True))
    checked = 0
    for r in range(R):
        order = np.argsort(-logits[r], kind="stable")  # logits order == probability order; ties -> lower index
        top = order[:K]
        # The reference sorts the fp32 PROBABILITIES (stable: equal probabilities -> lower index first, lrcn.jl:655-661).  Two distinct
        # logits closer than ~4e-7 can round to the same probability, and which pairs do depends on the last bit of the normaliser:
        # such rows (a handful among thousands of near-uniform ones) are compared as sets / skipped; exact logit ties stay strict.
        gaps = logits[r, order[:K]].astype(np.float64) - logits[r, order[1:K + 1]].astype(np.float64)
        ambiguous = (gaps > 0) & (gaps < 5e-7)
        if ambiguous[-1]:
            continue                                                # the K-th / (K+1)-th pair: membership itself is a coin toss
        if ambiguous.any():
            assert sorted(tok[r].tolist()) == sorted((top + 1).tolist()), f"row {r}"
            continue
        # bit-exact token ids whenever the logits around the cut are distinct (equal logits: the stable index rule decides)
        assert tok[r].tolist() == (top + 1).tolist(), f"row {r}"
        np.testing.assert_allclose(lp[r], lp_ref[r, top], rtol=1e-4, atol=2e-5)
        np.testing.assert_allclose(sc[r], np.exp(lp_ref[r, top]) * parent[r], rtol=5e-5)
        checked += 1
    assert checked >= 0.97 * R
    assert tok[16, :min(K, 3)].tolist() == [124, 4568, 10000][:min(K, 3)]
    assert tok[17].tolist() == list(range(1, K + 1))


def test_checkpoint_sidecar_roundtrip_through_the_library(tmp_path):
    """f-2: lrcn_checkpoint_save / _load (lrcn.jl:88-93,183-186,228-231): weights + Adam m, v, t + vocab, bit-exact; the
    pure-host reader (host.read_checkpoint, twin of the Julia reader) reads what the library wrote and vice versa."""
    from lrcn_b200 import host
    E, H1, H2, V, B, l = 32, 32, 64, 97, 6, 4
    model = synth.initweights([H1, H2], V, E, seed=3)
    feats = synth.features(8) * np.float32(50)
    ids = np.arange(1, 9, dtype=np.int64)
    tok = synth.tokens(l, B, V)
    vocab = {"~~": 1, "``": 2, "##": 3, "dog": 4, "café": 5}
    cfg = dict(embed=E, hidden1=H1, hidden2=H2, vocab=V, max_batch=B, max_len=l, max_gen_rows=8, precision=abi.PREC_FP32)
    p = str(tmp_path / "m.lrcnck")
    with abi.Handle(abi.default_config(**cfg)) as h:
        h.set_model(model)
        h.load_features(0, ids, feats)
        for _ in range(3):
            h.train_step(0, ids[:B], tok)
        w3 = h.get_model()
        m3 = [h.get_adam_state(k, 0) for k in range(1, 10)]
        v3 = [h.get_adam_state(k, 1) for k in range(1, 10)]
        h.checkpoint_save(p, True, host.vocab_to_bytes(vocab))
        L4 = h.train_step(0, ids[:B], tok)
        w4 = h.get_model()
    r = host.read_checkpoint(p)
    assert r["dims"] == (E, H1, H2, V) and r["adam_t"] == 3 and r["vocab"] == vocab
    for a, b in zip(w3 + m3 + v3, r["model"] + r["m"] + r["v"]):
        assert np.array_equal(a, b)
    with abi.Handle(abi.default_config(**cfg)) as h:   # resume: the 4th step after loading == the 4th step of the original run
        h.load_features(0, ids, feats)
        had, aux = h.checkpoint_load(p)
        assert had and host.vocab_from_bytes(aux) == vocab and h.get_adam_step() == 3
        for a, b in zip(w3, h.get_model()):
            assert np.array_equal(a, b)
        L4b = h.train_step(0, ids[:B], tok)
        assert abs(L4b - L4) < 1e-6 * abs(L4)
        for a, b in zip(w4, h.get_model()):
            np.testing.assert_allclose(a, b, rtol=1e-5, atol=1e-7)
        # a host-written model-only file loads too and leaves the optimizer alone; a wrong-shape file fails loudly
        host.write_checkpoint(p, model, (E, H1, H2, V), vocab=vocab)
        had, aux = h.checkpoint_load(p)
        assert not had and h.get_adam_step() == 4
        for a, b in zip(model, h.get_model()):
            assert np.array_equal(a, b)
        host.write_checkpoint(p, synth.initweights([H1, H2], V + 1, E), (E, H1, H2, V + 1))
        with pytest.raises(abi.LrcnError) as ei:
            h.checkpoint_load(p)
        assert ei.value.code == abi.ERR_ARG
