"""Portable synthetic-input generators for the LRCN decoder hot path.

No dataset files exist on the build or GPU boxes, so every input (weights, fc7
features, caption tokens, lengths, image ids) is produced by a counter-based
SplitMix64 stream that is trivially re-implementable in C++/Julia:

    z = seed*0x632BE59BD9B4E019 + (i+1)*0x9E3779B97F4A7C15          (mod 2^64)
    z = (z ^ (z>>30))*0xBF58476D1CE4E5B9 ; z = (z ^ (z>>27))*0x94D049BB133111EB
    z ^= z>>31 ;  u_i = (z>>40) * 2^-24   in [0,1)

Shapes/distributions follow SURVEY.md §8(d).  Weight shapes and init follow
`initweights` (reference lrcn.jl:489-510): xavier = uniform(+-sqrt(2/(rows+cols)))
drawn in fp64 then cast to fp32, biases 0 with the forget slice [0:H) = 1.
"""
from __future__ import annotations

import numpy as np

F_CNN = 4096  # reference lrcn.jl:28  `const cnnout = 4096`
EOS, BOS, UNK = 1, 2, 3  # reference lrcn.jl:248-255, tokenizer.jl:157-159

_G = np.uint64(0x9E3779B97F4A7C15)
_M1 = np.uint64(0xBF58476D1CE4E5B9)
_M2 = np.uint64(0x94D049BB133111EB)
_S = np.uint64(0x632BE59BD9B4E019)


def splitmix_u64(seed: int, n: int, offset: int = 0) -> np.ndarray:
    """n 64-bit outputs of the counter-based stream `seed`, starting at counter `offset`."""
    with np.errstate(over="ignore"):
        i = np.arange(offset + 1, offset + n + 1, dtype=np.uint64)
        z = np.uint64(seed) * _S + i * _G
        z = (z ^ (z >> np.uint64(30))) * _M1
        z = (z ^ (z >> np.uint64(27))) * _M2
        z = z ^ (z >> np.uint64(31))
    return z


def uniform01(seed: int, n: int, offset: int = 0) -> np.ndarray:
    """fp64 uniforms in [0,1) with 24 random bits (exactly representable in fp32)."""
    return (splitmix_u64(seed, n, offset) >> np.uint64(40)).astype(np.float64) * (2.0 ** -24)


def param_shapes(embed: int, hidden, vocab: int):
    """The frozen 9-matrix weight contract (reference lrcn.jl:489-510), 1-based order."""
    h1, h2 = int(hidden[0]), int(hidden[1])
    if h2 % 2:
        raise ValueError("hidden[2] must be even: lrcn() concatenates two ceil(H2/2) halves "
                         "into the H2-wide input of layer 2 (lrcn.jl:496-498,545-546)")
    c = (h2 + 1) // 2
    return [
        (embed + h1, 4 * h1),  # 1 W1
        (1, 4 * h1),           # 2 b1
        (h2 + h2, 4 * h2),     # 3 W2   (X = hidden[end] for k==2)
        (1, 4 * h2),           # 4 b2
        (h1, c),               # 5 Wf   = model[end-4]
        (F_CNN, c),            # 6 Wcnn = model[end-3]
        (vocab, embed),        # 7 Wemb = model[end-2]
        (h2, vocab),           # 8 Wout = model[end-1]
        (1, vocab),            # 9 bout = model[end]
    ]


def initweights(hidden, vocab: int, embed: int, seed: int = 1, dtype=np.float32):
    """Synthetic stand-in for `initweights(atype,hidden,vocab,embed)` (lrcn.jl:489-510).

    Returns a list of 9 Fortran-ordered (column-major, like Julia) arrays.
    """
    shapes = param_shapes(embed, hidden, vocab)
    model = []
    for k, (r, c) in enumerate(shapes):
        if r == 1:  # bias rows
            b = np.zeros((1, c), dtype=dtype, order="F")
            if k in (1, 3):  # model[2k][1:H] = 1  (forget gate bias)
                b[0, : c // 4] = 1
            model.append(b)
        else:
            s = np.sqrt(2.0 / (r + c))
            u = uniform01(seed * 16 + k, r * c)
            w = (2.0 * s * u - s).astype(dtype)
            model.append(np.asfortranarray(w.reshape((r, c), order="F")))
    return model


def eos_timed_model(hidden, vocab: int, embed: int, seed: int = 7, scale: float = 6.0, n_counters: int = 32,
                    w_eos: float = 0.12, g_bias: float = 0.12, v_scale: float = 0.012):
    """A synthetic model whose beam-search decodes END like real captions do ("bias bout[eos] or train").

    Random (untrained) weights never make eos the best continuation, so every decode would run the full nword+1 steps.
    Here `n_counters` units of layer 2 are turned into slow integrators (forget ~ 1, a steady positive change whose size
    depends on the image through the x_cnn half of the layer-2 input, no other inputs) and feed the eos logit, which
    therefore grows step by step and overtakes the other words after an image-dependent number of steps.  With the
    defaults at E = H = 512, V = 10000 (BASELINE configs[2]) beam-3 decode lengths spread over 2..32 tokens with a mean
    near the ~10.4 steps of COCO captions.  Everything else is initweights() scaled by `scale`."""
    h1, h2 = int(hidden[0]), int(hidden[1])
    c = (h2 + 1) // 2
    model = initweights(hidden, vocab, embed, seed=seed)
    model = [w * np.float32(scale) if w.shape[0] > 1 else w for w in model]
    W2, b2, Wout = model[2], model[3], model[7]
    u = np.arange(n_counters)
    for g in range(4):
        W2[:, g * h2 + u] = 0                      # counters see no q, no recurrent input ...
    b2[0, u] = 6.0                                 # forget ~ 1: the cell integrates
    b2[0, h2 + u] = 0.0                            # ingate = 0.5
    b2[0, 2 * h2 + u] = 3.0                        # outgate ~ 0.95
    b2[0, 3 * h2 + u] = g_bias                     # steady positive change
    col = (2.0 * uniform01(seed * 16 + 11, c) - 1.0) * 1.7 * v_scale
    W2[c:2 * c, 3 * h2:3 * h2 + n_counters] = col.astype(np.float32)[:, None]   # ... but the image (x_cnn) sets their rate
    Wout[u, EOS - 1] = w_eos                       # eos logit grows with the counters
    return model


def features(n_img: int, seed: int = 2) -> np.ndarray:
    """n_img x 4096 fp32, post-ReLU-like and L1-normalised like the reference's `featsn`
    (lrcn.jl:597 `input/sum(input)`). Row-major: one image's 4096 floats are contiguous."""
    u = uniform01(seed, n_img * F_CNN).reshape(n_img, F_CNN)
    f = np.maximum(0.0, 2.0 * u - 1.0)
    f /= f.sum(axis=1, keepdims=True)
    return np.ascontiguousarray(f.astype(np.float32))


def tokens(l: int, batch: int, vocab: int, seed: int = 3, zipf: bool = False) -> np.ndarray:
    """l x B int64, time-major, 1-based ids in [4, V] (never eos/bos; unk only if zipf)."""
    u = uniform01(seed, l * batch).reshape(l, batch)
    if not zipf:
        t = 4 + np.floor(u * (vocab - 3)).astype(np.int64)
        return np.minimum(t, vocab)
    # Zipf(1) on [4,V] by inverse CDF of 1/r, with ~1% unk: stresses scatter-add collisions
    n = vocab - 3
    r = np.floor(np.exp(u * np.log(n + 1.0))).astype(np.int64)  # in [1, n]
    t = 3 + np.clip(r, 1, n)
    unk = uniform01(seed + 1000003, l * batch).reshape(l, batch) < 0.01
    t[unk] = UNK
    return t


# caption-length histograms (words per caption, after tokenizer-style stripping) measured from the
# reference's eval/flickr_refs/* and eval/coco_refs/* (SURVEY.md §8d): value -> weight
_FLICKR_LEN = {5: 2, 6: 4, 7: 7, 8: 9, 9: 10, 10: 10, 11: 9, 12: 8, 13: 7, 14: 6, 15: 5, 16: 4, 17: 4,
               18: 3, 19: 3, 20: 2, 21: 2, 22: 1, 23: 1, 24: 1, 25: 1, 26: 0.5, 27: 0.3, 28: 0.2}
_COCO_LEN = {7: 3, 8: 14, 9: 22, 10: 21, 11: 15, 12: 10, 13: 6, 14: 4, 15: 2, 16: 1, 17: 0.7,
             18: 0.5, 19: 0.3, 20: 0.2, 22: 0.1, 25: 0.1, 28: 0.1}


def lengths(n_batches: int, shape: str = "flickr", seed: int = 4) -> np.ndarray:
    """One caption length l per batch (reference batches are equal-length, lrcn.jl:299-327)."""
    if shape == "fixed20":
        return np.full(n_batches, 20, dtype=np.int64)
    hist = _FLICKR_LEN if shape == "flickr" else _COCO_LEN
    vals = np.array(sorted(hist), dtype=np.int64)
    w = np.array([hist[int(v)] for v in vals], dtype=np.float64)
    cdf = np.cumsum(w) / w.sum()
    u = uniform01(seed, n_batches)
    return vals[np.minimum(np.searchsorted(cdf, u, side="right"), len(vals) - 1)]


def image_ids(batch: int, n_img: int, seed: int = 5) -> np.ndarray:
    """B int64 image ids, uniform with replacement over ids 1..n_img (id k -> feature row k-1)."""
    u = uniform01(seed, batch)
    return 1 + np.minimum(np.floor(u * n_img).astype(np.int64), n_img - 1)
