"""CPU oracle for the LRCN decoder hot path -- TEST INFRASTRUCTURE ONLY.

This file is a NumPy restatement of the reference's algorithm for the hot path
(ekinakyurek/Long-Term-Recurrent-Convolutional-NN, `lrcn.jl`).  It exists to CHECK the
CUDA path; it is never the thing measured or shipped.  Only `tests/`,
`__graft_entry__.smoke()` and the `cpu_baseline` / `--impl reference` legs of `bench.py`
may import it.  The product (`lrcn_b200`, `liblrcn_b200.so`) never does.

PARITY UNPINNED.  The reference ships no tests, golden vectors or fixtures for this path
(SURVEY.md §4, §8c), Julia/Knet are not installed in the image and cannot be installed
(no network), and the arithmetic lives in the un-vendored, un-pinned third-party package
Knet.jl (+AutoGrad.jl; Julia-0.5-era syntax in lrcn.jl => Knet ~0.8.x).  The Knet
primitives used by the path are therefore restated from their published semantics:
  logp(x,2)   : x - max_row(x) - log(sum_row(exp(x - max_row(x))))
  sigm(x)     : 1/(1+exp(-x))
  dropout(x,p): x .* (rand(size(x)) .> p) ./ (1-p); identity at p=0.  Knet's RNG stream is not reproducible, so the
               oracle takes the two masks of every step (lrcn.jl:542,547) EXPLICITLY (`masks=`): tests hand it the
               masks the CUDA path draws (dropout_masks() restates its counter hash) and compare at p=0.4, the
               reference's only training setting (lrcn.jl:227)
  xavier      : uniform(+-sqrt(2/(rows+cols)))
  Adam/update!: m=b1*m+(1-b1)*g; v=b2*v+(1-b2)*g.*g; w-=lr*(m/(1-b1^t))./(sqrt(v/(1-b2^t))+eps)
  x[idx,:]    : row gather; adjoint = dense zeros + add-at-index (accumulates repeats)
What pins this oracle instead (tests/test_oracle.py): known answers derived from the
reference's own semantics and slides (untrained loss == ln V = 8.953 / 9.272 of
presentation.pptx charts), hand-computed LSTM steps, fp64 finite differences, and an
independent torch-autograd evaluation of the same forward.

Conventions (the reference's): matrices are rows x cols with activations B x features;
token ids are 1-based Int64 on the API (eos=1,bos=2,unk=3, lrcn.jl:248-255).
All functions are dtype-generic: pass float32 params for the fp32 restatement
(what Knet computes) or float64 params for the arbitration shadow.
"""
from __future__ import annotations

import numpy as np

EOS, BOS, UNK = 1, 2, 3  # lrcn.jl:248-255


# --------------------------------------------------------------------------- primitives
def sigm(x):
    """Knet `sigm`."""
    one = x.dtype.type(1)
    return one / (one + np.exp(-x))


def logp(x):
    """Knet `logp(x,2)`: row-wise log-softmax (lrcn.jl:562)."""
    z = x - x.max(axis=1, keepdims=True)
    return z - np.log(np.exp(z).sum(axis=1, keepdims=True))


def initstate(model, batch):
    """lrcn.jl:512-526 minus the junk third layer: [h1,c1,h2,c2] zeros."""
    dt = model[0].dtype
    h1 = model[1].shape[1] // 4
    h2 = model[3].shape[1] // 4
    return [np.zeros((batch, h1), dt), np.zeros((batch, h1), dt),
            np.zeros((batch, h2), dt), np.zeros((batch, h2), dt)]


def lstm(weight, bias, hidden, cell, input):
    """lrcn.jl:528-538.  Gate column order is [forget, ingate, outgate, change]."""
    gates = np.hstack([input, hidden]) @ weight + bias
    hsize = hidden.shape[1]
    forget = sigm(gates[:, :hsize])
    ingate = sigm(gates[:, hsize:2 * hsize])
    outgate = sigm(gates[:, 2 * hsize:3 * hsize])
    change = np.tanh(gates[:, 3 * hsize:])
    cell = cell * forget + ingate * change
    hidden = outgate * np.tanh(cell)
    return hidden, cell


def lrcn(w, s, x_cnn, x_lstm):
    """lrcn.jl:540-551 at pdrop=0.  Mutates s[0..3] like the reference mutates s[1..4]."""
    x = x_lstm
    s[0], s[1] = lstm(w[0], w[1], s[0], s[1], x)
    x = s[0]
    x = x @ w[-5]
    x = np.hstack([x, x_cnn])
    s[2], s[3] = lstm(w[2], w[3], s[2], s[3], x)
    x = s[2]
    return x @ w[-2] + w[-1]


def _step_inputs_targets(sequence, rng, batch):
    """inputs [bos,w1..wl], targets [w1..wl,eos] (lrcn.jl:556,563-569,572-577); 0-based ids."""
    toks = [np.asarray(sequence[t], dtype=np.int64) for t in rng]
    bos = np.full(batch, BOS, np.int64)
    eos = np.full(batch, EOS, np.int64)
    ins = [bos] + toks
    tgt = toks + [eos]
    return [a - 1 for a in ins], [a - 1 for a in tgt]


def loss(param, state, input, sequence, rng):
    """lrcn.jl:553-581.  `sequence[t]` is a length-B vector of 1-based ids; `rng` the
    (0-based, Python) index range of this batch's time rows.  Returns a Python float
    (the reference's Float64 `-total/count`)."""
    batch = input.shape[0]
    state = [s.copy() for s in state]
    total = 0.0
    count = 0
    ins, tgt = _step_inputs_targets(sequence, rng, batch)
    x_cnn = input @ param[-4]
    rows = np.arange(batch)
    for u, y in zip(ins, tgt):
        lstm_input = param[-3][u, :]
        ypred = lrcn(param, state, x_cnn, lstm_input)
        ynorm = logp(ypred)
        total += float(ynorm[rows, y].sum())  # fp32 row gather + sum, accumulated in Float64
        count += batch
    return -total / count


def token_logps(param, state, input, sequence, rng):
    """Per-step, per-row log-prob of the target token (T x B); a finer-grained view of loss()."""
    batch = input.shape[0]
    state = [s.copy() for s in state]
    ins, tgt = _step_inputs_targets(sequence, rng, batch)
    x_cnn = input @ param[-4]
    rows = np.arange(batch)
    out = []
    for u, y in zip(ins, tgt):
        ynorm = logp(lrcn(param, state, x_cnn, param[-3][u, :]))
        out.append(ynorm[rows, y])
    return np.stack(out)


# --------------------------------------------------------------------------- gradient
def _cell_fwd(W, b, x, h_prev, c_prev):
    xh = np.hstack([x, h_prev])
    G = xh @ W + b
    H = h_prev.shape[1]
    f = sigm(G[:, :H]); i = sigm(G[:, H:2 * H]); o = sigm(G[:, 2 * H:3 * H]); g = np.tanh(G[:, 3 * H:])
    c = c_prev * f + i * g
    tc = np.tanh(c)
    h = o * tc
    return h, c, (xh, f, i, o, g, c_prev, tc)


def _cell_bwd(W, cache, dh, dc_next):
    xh, f, i, o, g, c_prev, tc = cache
    one = f.dtype.type(1)
    do = dh * tc
    dc = dc_next + dh * o * (one - tc * tc)
    df = dc * c_prev
    di = dc * g
    dg = dc * i
    dc_prev = dc * f
    dG = np.hstack([df * f * (one - f), di * i * (one - i), do * o * (one - o), dg * (one - g * g)])
    dW = xh.T @ dG
    db = dG.sum(axis=0, keepdims=True)
    dxh = dG @ W.T
    return dW, db, dxh, dc_prev


def drop_hash24(seed, site, idx):
    """The CUDA path's counter-based dropout hash (csrc/kernels_simt.cu drop_hash24), restated: SplitMix64 finaliser of
    (seed, site, element index) -> 24 random bits.  idx: uint64 array."""
    with np.errstate(over="ignore"):
        z = np.uint64(seed) + np.uint64(0x9E3779B97F4A7C15) * (idx.astype(np.uint64) + np.uint64(1)) \
            + np.uint64(0xD1B54A32D192ED03) * np.uint64(site + 1)
        z = (z ^ (z >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)
        z = (z ^ (z >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)
        z = z ^ (z >> np.uint64(31))
    return (z >> np.uint64(40)).astype(np.uint32)


def dropout_masks(pdrop, seed, T, B, E, C2, rank=0, dtype=np.float32):
    """The two dropout masks of every decoder step as the CUDA path draws them, already scaled by 1/(1-p) like Knet's
    dropout (lrcn.jl:542 on the word embedding, :547 on hcat(x*Wf, x_cnn)): (M0 [T][B][E], M1 [T][B][C2]).
    Element index = (t*B + i)*width + column; keep iff hash24 >= uint32(p * 2^24); the rank is mixed into the seed so that
    the shards of a data-parallel batch draw different masks."""
    p32 = np.float32(pdrop)
    thresh = np.uint32(int(float(p32) * 16777216.0))
    keep = dtype(np.float32(1.0) / (np.float32(1.0) - p32))
    with np.errstate(over="ignore"):
        s = np.uint64(seed) + np.uint64(0x9E3779B97F4A7C15) * np.uint64(rank)
    out = []
    for site, width in ((0, E), (1, C2)):
        idx = np.arange(T * B * width, dtype=np.uint64)
        m = np.where(drop_hash24(s, site, idx) >= thresh, keep, dtype(0)).astype(dtype)
        out.append(m.reshape(T, B, width))
    return out


def lossgradient(param, state, input, sequence, rng, masks=None):
    """What `lossgradient = grad(loss)` (lrcn.jl:583) returns: d loss / d param[k] for all 9
    tensors, hand-derived BPTT.  Also returns the loss.  h0/c0 constant.
    masks = (M0, M1): explicit, pre-scaled dropout masks of the two sites lrcn.jl:542,547 (None = pdrop 0)."""
    W1, b1, W2, b2, Wf, Wcnn, Wemb, Wout, bout = param
    dt = W1.dtype
    batch = input.shape[0]
    E = Wemb.shape[1]
    C = Wf.shape[1]
    ins, tgt = _step_inputs_targets(sequence, rng, batch)
    T = len(ins)
    rows = np.arange(batch)
    h1, c1, h2, c2 = [s.copy() for s in state]
    v = input @ Wcnn
    caches = []
    total = 0.0
    for t, (u, y) in enumerate(zip(ins, tgt)):
        e = Wemb[u, :]
        if masks is not None:
            e = e * masks[0][t]          # x = dropout(x_lstm, pdrop)            lrcn.jl:542
        h1, c1, k1 = _cell_fwd(W1, b1, e, h1, c1)
        q = h1 @ Wf
        z = np.hstack([q, v])
        if masks is not None:
            z = z * masks[1][t]          # x = dropout(hcat(x, x_cnn), pdrop)    lrcn.jl:547
        h2, c2, k2 = _cell_fwd(W2, b2, z, h2, c2)
        a = h2 @ Wout + bout
        lp = logp(a)
        total += float(lp[rows, y].sum())
        caches.append((u, y, k1, h1, k2, h2, lp))
    n_tok = batch * T
    g = [np.zeros_like(p) for p in param]
    dv = np.zeros_like(v)
    dh1r = np.zeros_like(h1); dc1 = np.zeros_like(c1)
    dh2r = np.zeros_like(h2); dc2 = np.zeros_like(c2)
    inv = dt.type(1.0 / n_tok) if dt == np.float32 else 1.0 / n_tok
    for t in range(T - 1, -1, -1):
        (u, y, k1, h1t, k2, h2t, lp) = caches[t]
        dA = np.exp(lp)
        dA[rows, y] -= dt.type(1)
        dA *= inv
        g[7] += h2t.T @ dA
        g[8] += dA.sum(axis=0, keepdims=True)
        dh2 = dA @ Wout.T + dh2r
        dW, db, dxh, dc2 = _cell_bwd(W2, k2, dh2, dc2)
        g[2] += dW; g[3] += db
        if masks is not None:
            dxh[:, :2 * C] *= masks[1][t]
        dq = dxh[:, :C]
        dv += dxh[:, C:2 * C]
        dh2r = dxh[:, 2 * C:]
        g[4] += h1t.T @ dq
        dh1 = dq @ Wf.T + dh1r
        dW, db, dxh, dc1 = _cell_bwd(W1, k1, dh1, dc1)
        g[0] += dW; g[1] += db
        de = dxh[:, :E]
        if masks is not None:
            de = de * masks[0][t]
        dh1r = dxh[:, E:]
        np.add.at(g[6], u, de)  # accumulates over repeated ids (bos row gets B contributions)
    g[5] = input.T @ dv
    return g, -total / n_tok


# --------------------------------------------------------------------------- optimiser
class Adam:
    """Knet `Adam()` defaults used by initparams (lrcn.jl:399-405): lr=1e-3, b1=.9, b2=.999,
    eps=1e-8, no clipping, per-parameter state (fstm, scndm, t)."""

    def __init__(self, lr=0.001, beta1=0.9, beta2=0.999, eps=1e-8):
        self.lr, self.beta1, self.beta2, self.eps = lr, beta1, beta2, eps
        self.t = 0
        self.fstm = None
        self.scndm = None


def initparams(model):
    return [Adam() for _ in model]


def update(param, gloss, optim):
    """Knet `update!(param,gloss,optim)` (call site lrcn.jl:394): dense Adam on all 9 tensors."""
    for w, g, p in zip(param, gloss, optim):
        dt = w.dtype.type
        if p.fstm is None:
            p.fstm = np.zeros_like(w)
            p.scndm = np.zeros_like(w)
        p.t += 1
        p.fstm *= dt(p.beta1); p.fstm += dt(1 - p.beta1) * g
        p.scndm *= dt(p.beta2); p.scndm += dt(1 - p.beta2) * (g * g)
        m_hat = p.fstm / dt(1 - p.beta1 ** p.t)
        v_hat = p.scndm / dt(1 - p.beta2 ** p.t)
        w -= dt(p.lr) * (m_hat / (np.sqrt(v_hat) + dt(p.eps)))


def train_step(param, optim, input, sequence, rng):
    """One iteration of the train1 hot loop (lrcn.jl:378,394): gradient, then Adam. Returns loss."""
    state = initstate(param, input.shape[0])
    g, l = lossgradient(param, state, input, sequence, rng)
    update(param, g, optim)
    return l


def average_loss(param, batches):
    """lrcn.jl:407-486 numerics: token-weighted mean NLL over batches, skipping l>28.
    `batches` = iterable of (input B x 4096, sequence list, rng)."""
    total = 0.0
    count = 0
    for input, sequence, rng in batches:
        l = len(rng)
        if l > 28:
            continue
        b = input.shape[0]
        n = b * (l + 1)
        total += -loss(param, initstate(param, b), input, sequence, rng) * n
        count += n
    return -total / count


# --------------------------------------------------------------------------- generation
def beam_search(x, states, input, param, nword, current=1, trace=None):
    """lrcn.jl:644-678, recursion unrolled into a loop (tail call => identical semantics).

    x: list of (tokens list[int 1-based], prob float32); states: list of [h1,c1,h2,c2].
    Scores are raw fp32 probability products; ties -> lower index (Julia sortperm is
    index-tie-broken); only beam 1 is expanded when current==1; finished beams are not frozen;
    stop iff the best kept hypothesis ends in eos or current>nword.
    `trace`, if a list, receives per-hypothesis lists of chosen-token log-probs (observational).
    """
    K = len(x)
    lps = [[] for _ in range(K)]
    while True:
        new_x = []
        new_lp = []
        for i in range(K):
            current_index = x[i][0][-1]
            current_probability = np.float32(x[i][1])
            lstm_input = param[-3][current_index - 1:current_index, :]
            ypred = lrcn(param, states[i], input, lstm_input)
            ln = logp(ypred).astype(np.float32).reshape(-1)
            ynorm = np.exp(ln)
            xmaxes = np.argsort(-ynorm, kind="stable")[:K]
            pmaxes = ynorm[xmaxes] * current_probability
            for j in range(K):
                new_x.append((x[i][0] + [int(xmaxes[j]) + 1], np.float32(pmaxes[j])))
                new_lp.append(lps[i] + [float(ln[xmaxes[j]])])
            if current == 1:
                break
        scores = np.array([c[1] for c in new_x], dtype=np.float32)
        sorted_ = np.argsort(-scores, kind="stable")
        xs = [new_x[k] for k in sorted_[:K]]
        lps = [new_lp[k] for k in sorted_[:K]]
        if xs[0][0][-1] == EOS or current > nword:
            if trace is not None:
                trace.extend(lps)
            return xs
        # parent of kept candidate r = ceil((listpos_r) / K), 1-based  == listpos0 // K
        states = [[a.copy() for a in states[int(sorted_[i]) // K]] for i in range(K)]
        x = xs
        current += 1


def generate(param, feat, nword, beam_width, trace=None):
    """Numeric part of lrcn.jl:585-642 for a precomputed 4096-d feature row.
    Returns (tokens incl. leading bos, fp32 path probability) of the best hypothesis."""
    feat = np.asarray(feat, dtype=param[0].dtype).reshape(1, -1)
    input = feat @ param[-4]
    state = initstate(param, 1)
    x = [([BOS], np.float32(1.0)) for _ in range(beam_width)]
    states = [[a.copy() for a in state] for _ in range(beam_width)]
    xs = beam_search(x, states, input, param, nword, 1, trace)
    return xs[0][0], xs[0][1]


def caption_text(word_indices, index_to_char):
    """Text formatting of lrcn.jl:633-640: tokens[2:] until first eos, each + ' ', then '.'."""
    out = []
    for tok in word_indices[1:]:
        if tok == EOS:
            break
        out.append(index_to_char[tok - 1] + " ")
    return "".join(out) + "."


# --------------------------------------------------------------------------- batching (f-1)
def delete_unbatchable_captions(lengths, batch_size):
    """Index-level restatement of lrcn.jl:299-327 on a length-sorted list of caption lengths.
    Returns the (0-based) indices that SURVIVE."""
    lengths = list(lengths)
    n = len(lengths)
    limit = n - batch_size + 1          # 1-based limit as in the reference
    max_length = max(lengths)
    current_length = lengths[0]
    current_index = 1                   # 1-based like the reference
    drop = []
    while current_index < limit:
        if lengths[current_index + batch_size - 2] == current_length:
            current_index += batch_size
        else:
            old_index = current_index
            current_index = 0
            while current_index == 0:
                current_length += 1
                if current_length > max_length:
                    break
                try:
                    current_index = lengths.index(current_length) + 1
                except ValueError:
                    current_index = 0
            if current_index == 0:
                # lrcn.jl:311-318: findfirst found no longer caption; the reference's outer loop then never advances
                raise RuntimeError("reference does not terminate on this input (unbatchable tail with no longer caption)")
            drop.extend(range(old_index, current_index))
        if current_index >= limit:
            drop.extend(range(current_index, n + 1))
            break
    dropset = set(drop)
    return [k - 1 for k in range(1, n + 1) if k not in dropset]
