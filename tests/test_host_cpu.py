"""Host-side mirror of the reference's epoch drivers (lrcn.jl:330-397 train1, :407-486 average_loss, :585-642 generate),
exercised on CPU against a recording stand-in for the C-ABI handle: which batches are visited, with which token rows,
image ids and options.  The numeric path itself is covered by the GPU parity tests."""
import io
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import lrcn_b200  # noqa: E402,F401
from lrcn_b200 import abi, host  # noqa: E402


class FakeHandle:
    def __init__(self, cfg):
        self.cfg = cfg
        self.calls = []

    def close(self):
        pass

    def sync(self):
        self.calls.append(("sync",))

    def checkpoint_save(self, path, with_adam, aux):
        self.calls.append(("save", path, with_adam, bytes(aux)))

    def train_step(self, split, ids, tok, pdrop, seed):
        self.calls.append(("train", split, np.array(ids), np.array(tok), pdrop, seed))
        return float(tok.shape[0])

    def train_epoch(self, split, sequence, input_ids, lengths, order, pdrop, seed):
        """What lrcn_train_epoch does with its arguments, restated: batch b owns lengths[b] rows of `sequence` from the prefix sum."""
        self.calls.append(("epoch", split, len(order)))
        starts = np.concatenate([[0], np.cumsum(lengths)])
        out = []
        for b in order:
            l = int(lengths[b])
            if l > 28:
                continue
            out.append(self.train_step(split, input_ids[b], sequence[starts[b]:starts[b] + l], pdrop, seed + len(out)))
        return out

    def loss_epoch(self, split, sequence, input_ids, lengths):
        starts = np.concatenate([[0], np.cumsum(lengths)])
        tot, cnt = 0.0, 0
        for b, l in enumerate(lengths):
            if l <= 28:
                s_, n_ = self.loss(split, input_ids[b], sequence[starts[b]:starts[b] + int(l)])
                tot, cnt = tot + s_, cnt + n_
        return tot, cnt

    def loss(self, split, ids, tok):
        self.calls.append(("loss", split, np.array(ids), np.array(tok)))
        l, B = tok.shape
        return -2.0 * B * (l + 1), B * (l + 1)  # every token costs 2 nats

    def beam_search(self, split, ids, K, nword):
        if int(ids[0]) == 404:
            raise abi.LrcnError(abi.ERR_MISSING, "no features for image 404")
        toks = np.array([[2, 5, 4, 1, 7, 0]], dtype=np.int64)
        return toks, np.array([5]), np.array([0.25], np.float32), np.zeros((1, 5), np.float32)


@pytest.fixture
def net(monkeypatch):
    monkeypatch.setattr(abi, "Handle", FakeHandle)
    return host.LRCN([8, 8], 20, 8, 2)


def make_seq(lengths_per_batch, B=2):
    """(sequence, input_ids, lengths) in the reference's format (lrcn.jl:257-297) for batches of the given caption lengths."""
    sequence, input_ids, lengths = [], [], []
    tok = 4
    for b, l in enumerate(lengths_per_batch):
        for _ in range(l):
            sequence.append(np.arange(tok, tok + B, dtype=np.int64))
            tok += B
        input_ids.append(np.arange(100 * b, 100 * b + B, dtype=np.int64))
        lengths += [l] * B
    return sequence, input_ids, lengths


def test_two_layers_only():
    with pytest.raises(ValueError):
        host.LRCN([8], 20, 8, 2)


@pytest.mark.parametrize("per_step", [False, True])
def test_train1_visits_every_batch_once_in_shuffled_order_and_skips_long_captions(net, per_step):
    per_batch = [3, 5, 29, 7, 28]  # 29 > 28 is skipped (lrcn.jl:353)
    seq = make_seq(per_batch)
    losses = net.train1(seq, pdrop=0.4, shuffle_seed=3, per_step=per_step)
    assert len([c for c in net.h.calls if c[0] == "epoch"]) == (0 if per_step else 1)  # default: one library call per epoch
    calls = [c for c in net.h.calls if c[0] == "train"]
    assert len(calls) == 4 and len(losses) == 4
    seen = sorted(c[3].shape[0] for c in calls)
    assert seen == [3, 5, 7, 28]
    starts = np.concatenate([[0], np.cumsum(per_batch)])
    for _, split, ids, tok, pdrop, seed in calls:
        l = tok.shape[0]
        b = per_batch.index(l)
        assert split == 0 and pdrop == 0.4
        assert np.array_equal(ids, seq[1][b])                                   # image ids of that batch
        assert np.array_equal(tok, np.stack(seq[0][starts[b]:starts[b] + l]))    # its l token rows, time-major
    assert [c[5] for c in calls] == [1, 2, 3, 4]                                # a fresh dropout seed per step
    net.train1(seq, pdrop=0.4, shuffle_seed=4, per_step=per_step)
    assert [c[5] for c in net.h.calls if c[0] == "train"][4:] == [5, 6, 7, 8]   # ... continuing over epochs
    again = host.LRCN([8, 8], 20, 8, 2)
    again.train1(seq, pdrop=0.4, shuffle_seed=3, per_step=not per_step)
    assert [c[3].shape[0] for c in again.h.calls if c[0] == "train"] == [c[3].shape[0] for c in calls]  # the shuffled order is a function of the seed


@pytest.mark.parametrize("per_step", [False, True])
def test_average_loss_is_token_weighted_and_skips_long_captions(net, per_step):
    seq = make_seq([3, 29, 6])
    val = net.average_loss(seq, per_step=per_step)
    calls = [c for c in net.h.calls if c[0] == "loss"]
    assert [c[3].shape[0] for c in calls] == [3, 6]
    assert abs(val - 2.0) < 1e-12  # -(sum of log-probs) / (number of tokens), pooled over batches (lrcn.jl:476-486)


def test_generate_prints_id_line_and_caption_like_the_reference(net):
    vocab = {f"w{k}": k for k in range(1, 21)}
    out, ids = io.StringIO(), io.StringIO()
    hyp = net.generate(17, vocab, 30, 3, out=out, in_out=ids)
    assert ids.getvalue() == "17\n"
    assert out.getvalue() == "w5 w4 .\n"       # tokens after bos up to the first eos, each + ' ', then '.' (lrcn.jl:633-640)
    assert hyp == [2, 5, 4, 1, 7]
    with pytest.raises(RuntimeError, match="misssing features"):  # sic, lrcn.jl:603
        net.generate(404, vocab, 30, 3, out=io.StringIO(), in_out=io.StringIO())


def test_delete_unbatchable_captions_leaves_whole_equal_length_batches():
    """Property of lrcn.jl:299-327 that minibatch (lrcn.jl:276-293) relies on: what survives is an order-preserving subset made
    of consecutive groups of batch_size captions of one length.  (A list no longer than batch_size passes through unchanged,
    as in the reference.)"""
    rs = np.random.RandomState(11)
    for _ in range(3000):
        bs = int(rs.randint(1, 6))
        n = int(rs.randint(1, 80))
        lengths = sorted(rs.randint(1, 8, size=n).tolist())
        caps = [((i, ["w"] * l), l) for i, l in enumerate(lengths)]
        out = host.delete_unbatchable_captions(caps, bs)
        ids = [t[0][0] for t in out]
        assert ids == sorted(ids) and set(ids) <= set(range(n))
        if n <= bs:
            assert len(out) == n
            continue
        L = [t[1] for t in out]
        assert len(L) % bs == 0
        assert all(len(set(L[i:i + bs])) == 1 for i in range(0, len(L), bs))
        # nothing batchable from the front of a length class is thrown away except by the reference's tail rule
        for length in set(lengths):
            assert L.count(length) <= lengths.count(length) - lengths.count(length) % bs


def test_delete_unbatchable_captions_equals_the_restated_reference_whenever_it_terminates():
    from oracle import lrcn_oracle as O
    rs = np.random.RandomState(5)
    compared = skipped = 0
    for _ in range(4000):
        bs = int(rs.randint(1, 6))
        n = int(rs.randint(bs + 1, 70))
        lengths = sorted(rs.randint(1, 9, size=n).tolist())
        caps = [((i, ["w"] * l), l) for i, l in enumerate(lengths)]
        try:
            keep = O.delete_unbatchable_captions(lengths, bs)
        except RuntimeError:
            skipped += 1  # lrcn.jl:311-318 loops forever here; the host mirror drops the tail instead
            continue
        got = [t[0][0] for t in host.delete_unbatchable_captions(caps, bs)]
        assert got == keep
        compared += 1
    assert compared > 1000


def test_train_loop_mirrors_the_reference_epoch_loop(net, tmp_path):
    """train! (lrcn.jl:222-239): per epoch train1 at pdrop 0.4, checkpoint, average_loss on both splits, one printed line."""
    trn, val = make_seq([3, 5, 4]), make_seq([2, 6])
    out = io.StringIO()
    sheet = tmp_path / "sheet.out"
    hist = net.train((trn, val), vocab={"a": 4}, epochs=2, savefile="ck.bin", out=out, datasheet=str(sheet))
    kinds = [c[0] for c in net.h.calls if c[0] in ("epoch", "save", "sync")]
    assert kinds == ["epoch", "save", "epoch", "save", "sync"]
    assert all(c[4] == 0.4 for c in net.h.calls if c[0] == "train")            # lrcn.jl:227 hard-codes pdrop = 0.4
    losses = [c for c in net.h.calls if c[0] == "loss"]
    assert [c[1] for c in losses] == [0, 0, 0, 1, 1] * 2                       # training split then validation split, every epoch
    lines = out.getvalue().splitlines()
    assert lines == ["(:epoch,1,:loss,2.0f0,2.0f0)", "(:epoch,2,:loss,2.0f0,2.0f0)"] and sheet.read_text().splitlines() == lines
    assert hist == [(2.0, 2.0), (2.0, 2.0)]
    orders = [[c[3].shape[0] for c in net.h.calls if c[0] == "train"][:3], [c[3].shape[0] for c in net.h.calls if c[0] == "train"][3:]]
    assert sorted(orders[0]) == sorted(orders[1]) == [3, 4, 5]                 # every batch once per epoch
