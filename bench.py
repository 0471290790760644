#!/usr/bin/env python
"""bench.py -- LRCN decoder hot path on B200 (contract in the task statement, tier section 4).

A "step" is one training step (forward, BPTT, [gradient allreduce], Adam) of the 2-layer factored LSTM
caption decoder on one synthetic batch.  Workload = BASELINE.json configs[1]: Flickr30k-shaped
(fc7 4096-d, E=H1=H2=512, V=7731), 256 captions per GPU per step, caption length l drawn per batch from
the Flickr length histogram (all ranks share l; rows differ per rank), data-parallel over N GPUs.

  value : whole-job tokens/s with the batches already resident in HBM (lrcn_train_step_staged)
  e2e   : the same metric through the reference-facing call lrcn_train_step with HOST buffers
          (token/id H2D and loss D2H inside the timed region)
  --impl reference : the CPU restatement of the Knet path (oracle/, numpy+OpenBLAS, all host threads) on a
          bounded sample of the same workload (Julia/Knet cannot run in this image: see DESIGN.md)
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

WORKLOADS = {
    # name: (E, H1, H2, V, B per GPU, length shape)
    "flickr30k_train_b256": dict(E=512, H1=512, H2=512, V=7731, B=256, shape="flickr"),
    "flickr8k_train_b64": dict(E=512, H1=512, H2=512, V=8000, B=64, shape="fixed20"),       # configs[0], Knet-CPU case
    "coco_2f_train_b256": dict(E=1000, H1=1000, H2=1000, V=10636, B=256, shape="coco"),     # configs[3]
}
N_IMG = 8192
N_SLOTS = 16


def flops_per_token(w):
    E, H1, H2, V = w["E"], w["H1"], w["H2"], w["V"]
    C = H2 // 2
    return 6.0 * ((E + H1) * 4 * H1 + H1 * C + 2 * H2 * 4 * H2 + H2 * V)  # 6*M_tok


class ClockSampler:
    """nvidia-smi clocks / throttle reasons DURING the timed region (B200_PROFILING.md recipe)."""
    Q = "clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown," \
        "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, gpu_index):
        self.rows, self.proc = [], None
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-i", str(gpu_index),
                                          "-lms", "100"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append((time.time(), line.strip()))

    def stop(self, t0, t1):
        if not self.proc:
            return None
        time.sleep(0.15)
        self.proc.terminate()
        sm, mx, reasons = [], 0.0, set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ts, line in self.rows:
            if ts < t0 - 0.05 or ts > t1 + 0.15:
                continue
            f = [x.strip() for x in line.split(",")]
            try:
                sm.append(float(f[0])); mx = max(mx, float(f[1]))
            except Exception:
                continue
            for nm, val in zip(names, f[3:7]):
                if val.lower().startswith("active"):
                    reasons.add(nm)
        if not sm:
            return None
        return {"sm_mhz": float(np.median(sm)), "sm_max_mhz": mx, "reasons": sorted(reasons), "samples": len(sm)}


def make_batches(w, rank, n_slots, seed=0):
    from lrcn_b200 import synth
    ls = synth.lengths(n_slots, w["shape"], seed=4 + seed)             # same on every rank
    out = []
    for s in range(n_slots):
        l = int(ls[s])
        tok = synth.tokens(l, w["B"], w["V"], seed=3 + 1000 * s + 77 * rank, zipf=True)
        img = synth.image_ids(w["B"], N_IMG, seed=5 + 1000 * s + 77 * rank)
        out.append((img, tok, l))
    return out


def run_reference(args, w, rank, world):
    """CPU restatement of the Knet path (oracle/), numpy + OpenBLAS on all host threads, bounded sample."""
    if rank != 0:
        return
    from lrcn_b200 import synth
    from oracle import lrcn_oracle as O
    cores = os.cpu_count() or 1
    rows = min(w["B"], 64)
    model = synth.initweights([w["H1"], w["H2"]], w["V"], w["E"], seed=1)
    opt = O.initparams(model)
    feats = synth.features(256, seed=2)
    batches = make_batches(dict(w, B=rows), 0, N_SLOTS)

    def step(i):
        img, tok, l = batches[i % N_SLOTS]
        X = feats[(img - 1) % 256]
        O.train_step(model, opt, X, list(tok), range(0, l))
        return rows * (l + 1)

    for i in range(args.warmup):
        step(i)
    t0 = time.perf_counter()
    ntok = 0
    for i in range(args.steps):
        ntok += step(args.warmup + i)
    dt = time.perf_counter() - t0
    val = ntok / dt
    sample = f"{args.steps} train steps on {rows}-row slices of the {w['B']}-row batches (same lengths, same model)"
    line = {"impl": "reference", "metric": "train tokens/s", "value": val, "unit": "tokens/s", "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": 1e3 * dt / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic", "config": {"workload": args.workload, **{k: w[k] for k in ("E", "H1", "H2", "V", "B")}},
            "cpu_baseline": {"value": val, "unit": "tokens/s", "cores": cores, "kind": "port", "sample": sample},
            "e2e": {"value": val, "unit": "tokens/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "note": "CPU restatement of lrcn.jl's Knet path (Julia/Knet not installable in this image)"}
    print(json.dumps(line), flush=True)


def cpu_baseline_leg(w):
    """Bounded oracle sample on the host cores (rank 0, N=1): ~10-30 s of CPU work."""
    from lrcn_b200 import synth
    from oracle import lrcn_oracle as O
    rows = min(w["B"], 64)
    model = synth.initweights([w["H1"], w["H2"]], w["V"], w["E"], seed=1)
    opt = O.initparams(model)
    feats = synth.features(256, seed=2)
    batches = make_batches(dict(w, B=rows), 0, 4)
    img, tok, l = batches[0]
    O.train_step(model, opt, feats[(img - 1) % 256], list(tok), range(0, l))  # warm-up (BLAS threads, page faults)
    t0 = time.perf_counter()
    ntok, n = 0, 0
    while time.perf_counter() - t0 < 12.0 and n < 12:
        img, tok, l = batches[n % 4]
        O.train_step(model, opt, feats[(img - 1) % 256], list(tok), range(0, l))
        ntok += rows * (l + 1)
        n += 1
    dt = time.perf_counter() - t0
    return {"value": ntok / dt, "unit": "tokens/s", "cores": os.cpu_count() or 1, "kind": "port",
            "sample": f"{n} oracle train steps (numpy/OpenBLAS fp32) on {rows}-row slices of the workload's batches"}


def beam_leg(local_rank, rank, world, prec, max_over_ranks, barrier, n_img=1024, K=3, nword=30):
    """BASELINE.json configs[2]: COCO-shaped beam-search generation (E=H=512, V=10000, beam 3, nword 30), images sharded
    across ranks with no collective.  Synthetic weights carry no caption structure, so eos never becomes the best
    continuation and every image decodes the full nword+1 = 31 steps: this is the WORST case per caption (real COCO captions
    stop after ~10.4 steps).  `row_steps_per_s` (beam rows advanced one step per second) is the length-independent
    figure.  Host ids in, host tokens out (e2e by construction); wall clock, max over ranks."""
    from lrcn_b200 import abi, synth
    E = H = 512
    V = 10000
    cfg = abi.default_config(embed=E, hidden1=H, hidden2=H, vocab=V, max_batch=8, max_len=2, max_gen_rows=n_img * K, device=local_rank,
                             precision=prec, use_graphs=0)
    with abi.Handle(cfg) as g:
        g.set_model(synth.initweights([H, H], V, E, seed=1))
        ids = np.arange(1, n_img + 1, dtype=np.int64) + 100000 * rank
        g.load_features(1, ids, synth.features(n_img, seed=6 + rank))
        g.beam_search(1, ids, K, nword, want_logps=False)  # warm-up
        barrier()
        t0 = time.perf_counter()
        reps = 3
        steps = 0
        for _ in range(reps):
            toks, lens, prob, _ = g.beam_search(1, ids, K, nword, want_logps=False)
            steps += int(lens.max()) - 1
        g.sync()
        dt = max_over_ranks(time.perf_counter() - t0)
    return {"metric": "beam-3 captions/s", "value": reps * n_img * world / dt, "unit": "captions/s", "images_per_gpu": n_img, "beam_width": K,
            "nword": nword, "vocab": V, "mean_len": float(lens.mean() - 1), "decode_steps": steps / reps, "ms_per_batch": 1e3 * dt / reps,
            "row_steps_per_s": n_img * K * steps * world / dt,
            "note": "synthetic weights never emit eos: every caption runs the full 31 steps (worst case; COCO captions stop after ~10.4)",
            "timing": "wall clock around lrcn_beam_search (host ids in, host tokens out), max over ranks"}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=60)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="flickr30k_train_b256", choices=sorted(WORKLOADS))
    ap.add_argument("--precision", default="bf16x3", choices=["bf16x3", "fp32"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-beam", action="store_true")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3)
    w = WORKLOADS[args.workload]
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))

    import lrcn_b200  # noqa: F401
    if args.impl == "reference":
        run_reference(args, w, rank, world)
        return

    from lrcn_b200 import abi, synth
    dist = None
    if world > 1:
        import torch
        import torch.distributed as dist
        torch.cuda.set_device(local_rank)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    prec = abi.PREC_BF16X3 if args.precision == "bf16x3" else abi.PREC_FP32
    cfg = abi.default_config(embed=w["E"], hidden1=w["H1"], hidden2=w["H2"], vocab=w["V"], max_batch=w["B"], max_len=28,
                             max_gen_rows=8, device=local_rank, precision=prec, use_graphs=1)
    h = abi.Handle(cfg)
    h.set_model(synth.initweights([w["H1"], w["H2"]], w["V"], w["E"], seed=1))
    h.load_features(0, np.arange(1, N_IMG + 1, dtype=np.int64), synth.features(N_IMG, seed=2))
    dp_exchange = None
    if world > 1:
        import torch
        uid = torch.zeros(abi.COMM_ID_BYTES, dtype=torch.uint8, device="cuda")
        if rank == 0:
            uid = torch.frombuffer(bytearray(abi.Handle.comm_unique_id()), dtype=torch.uint8).cuda()
        dist.broadcast(uid, 0)
        h.comm_init(bytes(uid.cpu().numpy().tobytes()), rank, world)
        # peer-memory gradient exchange (csrc/dp_p2p.cu): every rank exports its IPC blob, all-gather, import
        mine = torch.frombuffer(bytearray(h.p2p_export()), dtype=torch.uint8).cuda()
        blobs = [torch.zeros_like(mine) for _ in range(world)]
        dist.all_gather(blobs, mine)
        ok = 1
        try:
            h.p2p_import(b"".join(bytes(b.cpu().numpy().tobytes()) for b in blobs), rank, world)
        except abi.LrcnError as e:  # no peer access between these GPUs
            ok = 0
            print(f"[bench] rank {rank}: peer memory unavailable ({e})", file=sys.stderr)
        flag = torch.tensor([ok], dtype=torch.int32, device="cuda")
        dist.all_reduce(flag, op=dist.ReduceOp.MIN)  # the choice must be the same on every rank
        if int(flag.item()) == 0:
            os.environ["LRCN_DP_NCCL"] = "1"  # read by the library at the first data-parallel step
        dp_exchange = "nccl" if os.environ.get("LRCN_DP_NCCL") else "p2p"

    batches = make_batches(w, rank, N_SLOTS)
    for s, (img, tok, l) in enumerate(batches):
        h.stage_batch(s, 0, img, tok)

    def barrier():
        h.sync()
        if dist is not None:
            dist.barrier()
            import torch
            torch.cuda.synchronize()

    def max_over_ranks(x):
        if dist is None:
            return x
        import torch
        t = torch.tensor([x], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    # ---- setup: capture the CUDA graph of every step shape (one per distinct caption length) before anything is timed --
    # graph capture + instantiation is this framework's "compile" step and costs milliseconds per shape
    seen = set()
    for s_, (_, _, l_) in enumerate(batches):
        if l_ not in seen:
            seen.add(l_)
            h.train_step_staged(s_, 0.0, 7)
    h.sync()
    barrier()

    # ---- device-resident run: `value`
    for i in range(args.warmup):
        h.train_step_staged(i % N_SLOTS, 0.0, i)
    barrier()
    sampler = ClockSampler(local_rank) if rank == 0 else None
    launches0 = h.kernel_launches()
    t_wall0 = time.time()
    h.timer_start()
    ntok = 0
    for i in range(args.steps):
        s = (args.warmup + i) % N_SLOTS
        h.train_step_staged(s, 0.0, 1000 + i)
        ntok += w["B"] * (batches[s][2] + 1)
    ms = h.timer_stop()
    barrier()
    t_wall1 = time.time()
    launches = h.kernel_launches() - launches0
    clocks = sampler.stop(t_wall0, t_wall1) if sampler else None
    ms = max_over_ranks(ms)
    value = ntok * world / (ms * 1e-3)

    # ---- end-to-end run through the reference-facing call with host buffers: `e2e`
    for i in range(3):
        h.train_step(0, batches[i][0], batches[i][1], 0.0, i)
    barrier()
    h.timer_start()
    ntok_e = 0
    h2d = d2h = 0
    for i in range(args.steps):
        img, tok, l = batches[(args.warmup + i) % N_SLOTS]
        h.train_step(0, img, tok, 0.0, 2000 + i)          # H2D of tokens/ids/scalars and D2H of the loss inside
        ntok_e += w["B"] * (l + 1)
        h2d += (2 * (l + 1) * w["B"] + w["B"]) * 4 + 64
        d2h += 8
    ms_e = max_over_ranks(h.timer_stop())
    barrier()
    e2e_val = ntok_e * world / (ms_e * 1e-3)

    # ---- secondary metric of BASELINE.json: beam-3 captions/s (COCO-shaped generation, images sharded by rank, no collective)
    beam = None
    if not args.no_beam:
        beam = beam_leg(local_rank, rank, world, prec, max_over_ranks, barrier)

    # leave valid activations/gradients in the buffers for the single-kernel timings below (collective: every rank steps)
    h.train_step_staged(0, 0.0, 1)
    barrier()

    line = None
    if rank == 0:
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except Exception:
            pass
        peak_tf = peaks.get("bf16_tflops", 1590.0)
        peak_bw = peaks.get("hbm_gbs", 6650.0)
        src = "measured (MEASURED_PEAKS.json, burst)" if peaks else "fallback (B200_PROFILING.md)"
        # dominant kernel: the vocab-projection GEMM (49% of MACs/token); timed alone, L2 flushed between launches
        k_ms, k_bytes, k_flops = h.time_kernel("vocab_gemm", 10)
        a_ms, a_bytes, _ = h.time_kernel("adam", 10)
        ncu = {}
        try:
            ncu = json.load(open(os.path.join(ROOT, "profiles", "ncu_traffic.json"))).get(args.workload, {})
        except Exception:
            pass
        roof = {"bound": "tensor", "kernel": "gemm2_bf16x3_kernel<K,K> (vocab projection h2*Wout+bout: 2-CTA 256x256 tcgen05, 3 split passes)" if prec else "sgemm_kernel",
                "achieved": k_flops / (k_ms * 1e-3) / 1e12, "peak": peak_tf, "unit": "TFLOP/s",
                "frac": k_flops / (k_ms * 1e-3) / 1e12 / peak_tf,
                "traffic": ncu.get("vocab_gemm", {}).get("traffic_bytes") if prec else None, "peak_source": src,
                "algorithmic_flops": k_flops, "algorithmic_bytes": k_bytes, "launch_ms": k_ms,
                "tensor_pipe_active_pct_ncu": ncu.get("vocab_gemm", {}).get("tensor_pipe_active_pct") if prec else None,
                "note": "achieved counts ALGORITHMIC flops (2MNK); the bf16x3 split issues 3x that on the tensor pipe, so the ceiling of frac is 1/3; "
                        "L2 flushed between the timed launches"}
        roof_adam = {"bound": "hbm", "kernel": "adam_kernel", "achieved": a_bytes / (a_ms * 1e-3) / 1e9, "peak": peak_bw, "unit": "GB/s",
                     "frac": a_bytes / (a_ms * 1e-3) / 1e9 / peak_bw, "traffic": ncu.get("adam", {}).get("traffic_bytes") if prec else None,
                     "algorithmic_bytes": a_bytes, "launch_ms": a_ms, "peak_source": src}
        step_flops = flops_per_token(w) * ntok / args.steps
        line = {"metric": "train tokens/s", "value": value, "unit": "tokens/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
                "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
                "dtype": "bf16x3 (bf16 hi/lo split on tcgen05, fp32 accumulate; fp32-equivalent)" if prec else "f32", "data": "synthetic",
                "config": {"workload": args.workload, **{k: w[k] for k in ("E", "H1", "H2", "V")}, "batch_per_gpu": w["B"],
                           "global_batch": w["B"] * world, "lengths": w["shape"], "parallelism": f"dp{world}" + (f" ({dp_exchange} gradient exchange)" if dp_exchange else ""),
                           "l2": "per-step working set (params+grads+Adam 4x53 MB, logits ~100 MB) exceeds the 126 MB L2; no explicit flush"},
                "e2e": {"value": e2e_val, "unit": "tokens/s", "h2d_bytes_per_step": h2d // args.steps, "d2h_bytes_per_step": d2h // args.steps,
                        "ms_per_step": ms_e / args.steps},
                "gpu_launches": launches, "clocks": clocks,
                "step_tflops_algorithmic": step_flops / (ms / args.steps * 1e-3) / 1e12,
                "roofline": roof, "roofline_adam": roof_adam}
        if beam is not None:
            line["beam"] = beam
        if world == 1 and not args.no_cpu_baseline:
            line["cpu_baseline"] = cpu_baseline_leg(w)
    if dist is not None:
        dist.barrier()  # nobody unmaps its arenas while a peer may still be inside an exchange kernel
    h.close()
    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()
    if line is not None:
        print(json.dumps(line), flush=True)


if __name__ == "__main__":
    main()
