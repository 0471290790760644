"""Data-parallel invariance on real GPUs (run under torchrun, one rank per GPU):
an N-rank gradient/Adam step on a global batch == the oracle's single-device step on the same batch."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import lrcn_b200  # noqa: E402,F401
from lrcn_b200 import abi, synth  # noqa: E402
from oracle import lrcn_oracle as O  # noqa: E402


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    E, H1, H2, V, b, l = 64, 64, 64, 300, 8, 5
    Bg = b * world
    model = synth.initweights([H1, H2], V, E, seed=1)
    model = [w * np.float32(3) if w.shape[0] > 1 else w for w in model]
    feats = synth.features(32, seed=2) * np.float32(50)
    ids = np.arange(1, 33, dtype=np.int64)
    img = synth.image_ids(Bg, 32)
    tok = synth.tokens(l, Bg, V)
    ok = True
    for prec in (abi.PREC_FP32, abi.PREC_BF16X3):
        cfg = abi.default_config(embed=E, hidden1=H1, hidden2=H2, vocab=V, max_batch=b, max_len=8, max_gen_rows=8, device=local, precision=prec)
        h = abi.Handle(cfg)
        h.set_model(model)
        h.load_features(0, ids, feats)
        uid = torch.zeros(abi.COMM_ID_BYTES, dtype=torch.uint8, device="cuda")
        if rank == 0:
            uid = torch.frombuffer(bytearray(abi.Handle.comm_unique_id()), dtype=torch.uint8).cuda()
        dist.broadcast(uid, 0)
        h.comm_init(uid.cpu().numpy().tobytes(), rank, world)
        mine = torch.frombuffer(bytearray(h.p2p_export()), dtype=torch.uint8).cuda()
        blobs = [torch.zeros_like(mine) for _ in range(world)]
        dist.all_gather(blobs, mine)
        h.p2p_import(b"".join(bytes(x.cpu().numpy().tobytes()) for x in blobs), rank, world)
        sl = slice(rank * b, (rank + 1) * b)
        L = h.grad(0, img[sl], np.ascontiguousarray(tok[:, sl]))
        g = [h.get_grad(k) for k in range(1, 10)]
        L2 = h.train_step(0, img[sl], np.ascontiguousarray(tok[:, sl]))
        w_after = h.get_model()
        for it in range(3):  # a few more steps: barrier epochs, graph replays
            h.train_step(0, img[sl], np.ascontiguousarray(tok[:, sl]))
        # Adam state after 4 steps (peer-memory mode keeps m, v sharded by owner: get_adam_state gathers them)
        dist.barrier()
        m_got = [h.get_adam_state(k, 0) for k in range(1, 10)]
        v_got = [h.get_adam_state(k, 1) for k in range(1, 10)]
        dist.barrier()
        final = np.concatenate([x.ravel() for x in h.get_model()])
        digest = torch.tensor([float(np.abs(final).sum()), float((final * np.arange(1, final.size + 1, dtype=np.float64) % 7).sum())], dtype=torch.float64, device="cuda")
        digests = [torch.zeros_like(digest) for _ in range(world)]
        dist.all_gather(digests, digest)
        same = all(bool(torch.equal(d, digests[0])) for d in digests)
        if rank == 0:
            g_ref, L_ref = O.lossgradient(model, O.initstate(model, Bg), feats[img - 1], list(tok), range(0, l))
            errs = [float(np.linalg.norm(g[k] - g_ref[k]) / np.linalg.norm(g_ref[k])) for k in range(9)]
            ref = [w.copy() for w in model]
            O.update(ref, g_ref, O.initparams(ref))
            derr = [float(np.linalg.norm((w_after[k] - model[k]) - (ref[k] - model[k])) / (np.linalg.norm(ref[k] - model[k]) + 1e-30)) for k in range(9)]
            ref4 = [w.copy() for w in model]
            opt4 = O.initparams(ref4)
            for _ in range(4):
                g4, _ = O.lossgradient(ref4, O.initstate(ref4, Bg), feats[img - 1], list(tok), range(0, l))
                O.update(ref4, g4, opt4)
            merr = max(float(np.linalg.norm(m_got[k] - opt4[k].fstm) / (np.linalg.norm(opt4[k].fstm) + 1e-30)) for k in range(9))
            verr = max(float(np.linalg.norm(v_got[k] - opt4[k].scndm) / (np.linalg.norm(opt4[k].scndm) + 1e-30)) for k in range(9))
            if merr >= 2e-3 or verr >= 2e-3:
                for k in range(9):
                    a_, b_ = m_got[k].ravel(), opt4[k].fstm.ravel()
                    bad = np.nonzero(np.abs(a_ - b_) > 1e-3 * np.abs(b_).max())[0]
                    print(f"  tensor {k + 1} shape {m_got[k].shape}: m relerr {np.linalg.norm(a_ - b_) / (np.linalg.norm(b_) + 1e-30):.2e}, "
                          f"{bad.size} bad of {a_.size}, first bad {bad[:3]}, got {a_[bad[:3]]}, want {b_[bad[:3]]}", flush=True)
            same = same and merr < 2e-3 and verr < 2e-3
            good = abs(L - L_ref) < 1e-4 * abs(L_ref) and max(errs) < 1e-4 and abs(L2 - L_ref) < 1e-4 * abs(L_ref) and max(derr) < 5e-3 and same
            ok = ok and good
            print(f"dp_check world={world} prec={prec}: loss {L:.6f} vs {L_ref:.6f}; max grad relerr {max(errs):.2e}; max update relerr {max(derr):.2e}; replicas identical + Adam state (m {merr:.1e}, v {verr:.1e}) after 4 steps: {same} -> {'OK' if good else 'FAIL'}", flush=True)
        dist.barrier()
        h.close()
    dist.barrier()
    dist.destroy_process_group()
    sys.exit(0 if ok else 1)


if __name__ == "__main__":
    main()
