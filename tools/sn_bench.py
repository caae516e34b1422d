"""Spectral-norm / Adam micro-benchmark: achieved HBM GB/s of the grouped power iteration + scale (and backward) for the
BigGAN-128 networks.  Algorithmic bytes per parameter: forward 4 (one fp32 read for both matvecs) + 4 (fp32 read of the
scale pass) + 2 (bf16 normalised copy) [+ 2 transposed copy]; backward 8 (dot: G, W) + 8 (apply: G read, dW write).
Usage (B200):  python tools/sn_bench.py [--iters 20]"""
import argparse
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

ap = argparse.ArgumentParser()
ap.add_argument("--iters", type=int, default=20)
ap.add_argument("--config", type=int, default=1)
args = ap.parse_args()

from biggan_tensorflow_b200 import model, step

kw = dict(model.CONFIGS[args.config])
kw["batch_size"] = 8
cfg = model.Config(**kw)
tr = step.Trainer(cfg, compute_dtype=torch.bfloat16)
tr.build()
for net, grp in tr.sn_groups.items():
    n = sum(int(it.K) * int(it.C) for it in grp.items)
    nt = sum(int(it.K) * int(it.C) for it in grp.items if it.out_t)
    b16 = sum(int(it.K) * int(it.C) for it in grp.items if it.dtype != 0)
    fwd_bytes = 8 * n + 2 * b16 + 4 * (n - b16) + 2 * nt
    bwd_bytes = 16 * n
    for name, fn, nbytes in (("forward", lambda: grp.forward(tr.store), fwd_bytes), ("backward", grp.backward, bwd_bytes)):
        for _ in range(3):
            fn()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(args.iters):
            fn()
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / args.iters
        print("%s sn_group_%s: %d matrices, %.1f M params, %.3f ms, %.0f GB/s algorithmic" %
              (net, name, len(grp.items), n / 1e6, ms, nbytes / ms / 1e6))
