"""Fused attention core (bg_attn2_fwd / bg_attn2_bwd) alone through the C ABI at the BigGAN-128 shapes: ms and TFLOP/s.
Usage (B200):  python tools/attn_bench.py [--which g|d|both] [--op fwd|bwd|both] [--B 256] [--iters 20]
  g: Nq 4096, Nk 1024, dq 32 (24 padded), dv 96;  d: dq 16 (12 padded), dv 48.
Run under ncu with --iters 1 --warmup 1 for a capture of exactly these kernels."""
import argparse
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

ap = argparse.ArgumentParser()
ap.add_argument("--which", default="both", choices=["g", "d", "both"])
ap.add_argument("--op", default="both", choices=["fwd", "bwd", "both"])
ap.add_argument("--B", type=int, default=256)
ap.add_argument("--iters", type=int, default=20)
ap.add_argument("--warmup", type=int, default=3)
args = ap.parse_args()

from biggan_tensorflow_b200._lib import call, query
from biggan_tensorflow_b200.tensor import stream

dev = torch.device("cuda", 0)


def timed(fn):
    for _ in range(args.warmup):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / args.iters


for name, dq, dv in (("g", 32, 96), ("d", 16, 48)):
    if args.which not in (name, "both"):
        continue
    B, Nq, Nk = args.B, 4096, 1024
    q = torch.randn(B, Nq, dq, device=dev).mul_(0.3).bfloat16()
    k = torch.randn(B, Nk, dq, device=dev).mul_(0.3).bfloat16()
    v = torch.randn(B, Nk, dv, device=dev).bfloat16()
    o = torch.empty(B, Nq, dv, device=dev, dtype=torch.bfloat16)
    lse = torch.empty(B, Nq, device=dev, dtype=torch.float32)
    fwd = lambda: call("bg_attn2_fwd", q, k, v, o, lse, B, Nq, Nk, dq, dv, stream())
    nbf = int(query("bg_attn2_fwd_workspace_bytes", B))
    wsf = torch.empty(nbf, dtype=torch.uint8, device=dev)
    fwd_ws = lambda: call("bg_attn2_fwd_ws", q, k, v, o, lse, B, Nq, Nk, dq, dv, wsf, nbf, stream())
    if args.op in ("fwd", "both"):
        for label, f in (("two-pass", fwd), ("score-bound (single pass if bound <= 30)", fwd_ws), ("two-pass", fwd), ("score-bound", fwd_ws)):
            ms = timed(f)
            print("attn2_fwd %s b%d dq%d dv%d %s: %.3f ms  %.1f TFLOP/s" % (name, B, dq, dv, label, ms, 2.0 * B * Nq * Nk * (dq + dv) / ms / 1e9),
                  flush=True)
        print("  max bound %.2f" % float(wsf[:4 * B].view(torch.float32).max()))
    if args.op in ("bwd", "both"):
        fwd()
        do = torch.randn_like(o)
        gq, gk, gv = torch.empty_like(q), torch.empty_like(k), torch.empty_like(v)
        delta = torch.empty(B, Nq, device=dev, dtype=torch.float32)
        nb = int(query("bg_attn2_bwd_workspace_bytes", B, Nq, Nk, dq))
        ws = torch.empty(max(nb, 16), dtype=torch.uint8, device=dev)
        bwd = lambda: call("bg_attn2_bwd", q, k, v, o, do, lse, delta, gq, gk, gv, B, Nq, Nk, dq, dv, ws if nb else None, nb, stream())
        ms = timed(bwd)
        print("attn2_bwd %s b%d dq%d dv%d: %.3f ms  %.1f TFLOP/s" % (name, B, dq, dv, ms, 2.0 * B * Nq * Nk * (3 * dq + 4 * dv) / ms / 1e9), flush=True)
