"""N>1 host-side logic on CPU: two gloo ranks shard a global batch the way the GPU data-parallel path does
(host.shard_batch), compute their shard gradients with the oracle, all-reduce, and must reproduce the
single-process gradient (DP invariance).  Also: bench batches share lengths across ranks."""
import os
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import lrcn_b200  # noqa: F401
from lrcn_b200 import host, synth
from oracle import lrcn_oracle as O


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, out):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    torch.set_num_threads(1)
    E, H1, H2, V, Bg, l = 8, 12, 16, 40, 6, 4
    model = [w.astype(np.float64) * 3 if w.shape[0] > 1 else w.astype(np.float64) for w in synth.initweights([H1, H2], V, E, seed=1)]
    feats = synth.features(16, seed=2).astype(np.float64) * 50
    img = synth.image_ids(Bg, 16)
    tok = synth.tokens(l, Bg, V)
    my_img, my_tok = host.shard_batch(img, tok, rank, world)
    g, L = O.lossgradient(model, O.initstate(model, len(my_img)), feats[my_img - 1], list(my_tok), range(0, l))
    # the library scales by 1/(world*b*T); the oracle scaled by 1/(b*T): rescale, then allreduce(sum)
    flat = torch.from_numpy(np.concatenate([x.ravel(order="F") for x in g]) / world)
    loss = torch.tensor([L / world], dtype=torch.float64)
    dist.all_reduce(flat)
    dist.all_reduce(loss)
    g_full, L_full = O.lossgradient(model, O.initstate(model, Bg), feats[img - 1], list(tok), range(0, l))
    ref = np.concatenate([x.ravel(order="F") for x in g_full])
    err = float(np.abs(flat.numpy() - ref).max() / np.abs(ref).max())
    ids, (lo, hi) = host.shard_images(np.arange(10), rank, world)
    out[rank] = (err, abs(float(loss) - L_full), lo, hi)
    dist.destroy_process_group()


def test_two_rank_sharded_gradient_equals_global_gradient():
    world = 2
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_worker, args=(world, _free_port(), out), nprocs=world, join=True)
    assert len(out) == world
    for r in range(world):
        err, dl, lo, hi = out[r]
        assert err < 1e-12 and dl < 1e-12
    assert (out[0][2], out[0][3], out[1][2], out[1][3]) == (0, 5, 5, 10)


def test_bench_batches_share_lengths_across_ranks():
    import bench
    w = dict(bench.WORKLOADS["flickr30k_train_b256"], B=8)
    a = bench.make_batches(w, 0, 6)
    b = bench.make_batches(w, 1, 6)
    assert [x[2] for x in a] == [x[2] for x in b]          # same caption length per step on every rank
    assert any((x[1] != y[1]).any() for x, y in zip(a, b))  # different rows
    assert all(x[1].shape == (x[2], 8) and x[1].min() >= 3 for x in a)


def test_shard_batch_rejects_ragged_split():
    import pytest
    with pytest.raises(ValueError):
        host.shard_batch(np.arange(5), np.zeros((2, 5), np.int64), 0, 2)
