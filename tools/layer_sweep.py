"""Several conv layers of the BigGAN-128 step through the C ABI in ONE process (bf16 tensor-core path): ms and TFLOP/s per
layer, for A/B runs of two builds inside one GPU job (tools/layer_bench.py times a single geometry).
Usage (B200):  python tools/layer_sweep.py [--set short|deep|all] [--iters 20] [--json out.json]
Spec: (op, B, H, Ci, Co, k, stride, deconv, pad, residual, relu_mask)."""
import argparse
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

SHORT = [  # layers with a short main loop (epilogue / ingest bound)
    ("fwd", 512, 64, 48, 96, 1, 1, 0, "zero", 1, 0),      # D attn_conv (+x residual)
    ("fwd", 256, 64, 96, 192, 1, 1, 0, "zero", 1, 0),     # G attn_conv (+x residual)
    ("fwd", 256, 64, 192, 160, 1, 1, 0, "zero", 0, 0),    # G stacked f/g/h projection
    ("fwd", 512, 64, 32, 96, 1, 1, 0, "zero", 0, 0),      # im2col RGB GEMM of D's first convs
    ("fwd", 256, 128, 96, 96, 3, 1, 1, "zero", 1, 0),     # G last block, second conv (halo kernel, + skip residual)
    ("fwd", 256, 64, 192, 96, 4, 2, 1, "zero", 0, 0),     # G last block, 4x4/s2 transposed conv
    ("fwd", 512, 64, 96, 96, 3, 1, 0, "reflect", 1, 0),   # D block 1, second conv
    ("fwd", 512, 64, 96, 192, 3, 2, 0, "reflect", 0, 0),  # D block 2, strided conv
    ("dgrad", 512, 64, 96, 96, 3, 1, 0, "reflect", 0, 1),  # fold-mode dgrad + ReLU mask
    ("dgrad", 512, 64, 96, 192, 3, 2, 0, "reflect", 1, 1),  # fold-mode dgrad of the strided conv (+ accumulate)
    ("dgrad", 256, 128, 96, 16, 3, 1, 0, "reflect", 0, 0),  # RGB layer dgrad
    ("fwd", 256, 128, 96, 16, 3, 1, 0, "reflect", 0, 0),   # RGB layer forward (halo kernel, N = 16)
    ("dgrad", 256, 64, 192, 96, 4, 2, 1, "zero", 0, 0),    # dgrad of the 4x4/s2 transposed conv (a strided conv)
    ("wgrad", 256, 128, 96, 96, 3, 1, 1, "zero", 0, 0),
]
DEEP = [
    ("dgrad", 512, 4, 1536, 1536, 3, 1, 0, "reflect", 0, 1),
    ("dgrad", 256, 4, 1536, 1536, 3, 1, 0, "reflect", 1, 1),
    ("dgrad", 512, 8, 768, 768, 3, 1, 0, "reflect", 0, 1),
    ("fwd", 256, 8, 1536, 1536, 3, 1, 1, "zero", 0, 0),
    ("fwd", 256, 16, 768, 384, 4, 2, 1, "zero", 0, 0),
    ("fwd", 256, 32, 384, 384, 3, 1, 1, "zero", 1, 0),
    ("wgrad", 256, 16, 768, 768, 3, 1, 1, "zero", 0, 0),
]

ap = argparse.ArgumentParser()
ap.add_argument("--set", default="all", choices=["short", "deep", "all"])
ap.add_argument("--iters", type=int, default=20)
ap.add_argument("--warmup", type=int, default=3)
ap.add_argument("--only", type=int, default=-1, help="run only this index of the chosen set (ncu captures)")
ap.add_argument("--json", default="")
ap.add_argument("--variants", default="", help="';'-separated env settings timed INTERLEAVED per layer, e.g. "
                "'BG_TC_AUX_BATCH=0;BG_TC_AUX_BATCH=1;BG_TC_AUX_BATCH=1,BG_TC_AUX_PREFETCH=1' (knobs the library reads per call)")
ap.add_argument("--rounds", type=int, default=3, help="timing rounds per variant (the minimum is reported)")
args = ap.parse_args()

from biggan_tensorflow_b200 import _lib
from biggan_tensorflow_b200._lib import ConvDesc, call, query
from biggan_tensorflow_b200.tensor import stream

dev = torch.device("cuda", 0)
dt = _lib.BG_BF16


def run(spec):
    op, B, H, Ci, Co, k, s, deconv, pad, residual, relu_mask = spec
    if deconv:
        Ho = H * s
        tot = max((H - 1) * s + k - Ho, 0)
        p = tot // 2
        d = ConvDesc(B, H, H, Ci, Ho, Ho, Co, k, k, s, p, p, tot - p, tot - p, _lib.PAD_ZERO, 1)
        w = torch.randn(k, k, Co, Ci, device=dev).mul_(0.05).bfloat16()
    else:
        pd = (k - 1) // 2
        tot = 2 * pd if H % s == 0 else max(k - H % s, 0)
        pt = tot // 2
        Ho = (H + tot - k) // s + 1
        d = ConvDesc(B, H, H, Ci, Ho, Ho, Co, k, k, s, pt, pt, tot - pt, tot - pt,
                     _lib.PAD_REFLECT if (pad == "reflect" and tot) else _lib.PAD_ZERO, 0)
        w = torch.randn(k, k, Ci, Co, device=dev).mul_(0.05).bfloat16()
    w_t = w.transpose(2, 3).contiguous()
    x = torch.randn(B, H, H, Ci, device=dev).bfloat16()
    y = torch.empty(B, Ho, Ho, Co, device=dev, dtype=torch.bfloat16)
    flops = 2.0 * B * (H * H if deconv else Ho * Ho) * k * k * Ci * Co

    def ws_for(name):
        n = int(query(name, d, dt))
        return torch.empty(max(n, 16), dtype=torch.uint8, device=dev), n

    if op == "fwd":
        ws, n = ws_for("bg_conv2d_fwd_workspace_bytes")
        res = torch.randn_like(y) if residual else None
        fn = lambda: call("bg_conv2d_fwd", d, x, w, w_t, None, res, y, _lib.ACT_NONE, dt, ws if n else None, n, 0, stream())
    elif op == "dgrad":
        dy = torch.randn(B, Ho, Ho, Co, device=dev).bfloat16()
        dx = torch.empty_like(x)
        ws, n = ws_for("bg_conv2d_dgrad_workspace_bytes")
        res = torch.randn_like(dx) if residual else None
        mask = torch.randn_like(dx) if relu_mask else None
        fn = lambda: call("bg_conv2d_dgrad", d, dy, w, w_t, res, mask, dx, dt, ws if n else None, n, 0, stream())
    else:
        dy = torch.randn(B, Ho, Ho, Co, device=dev).bfloat16()
        dw = torch.empty(w.shape, device=dev, dtype=torch.float32)
        ws, n = ws_for("bg_conv2d_wgrad_workspace_bytes")
        fn = lambda: call("bg_conv2d_wgrad", d, x, dy, dw, dt, ws, ws.numel(), 0, stream())
    def timed():
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

    vms = []
    if args.variants:
        variants = [v for v in args.variants.split(";")]
        best = [1e30] * len(variants)
        for _ in range(args.rounds):
            for vi, v in enumerate(variants):
                saved = {}
                for kv in filter(None, v.split(",")):
                    key, val = kv.split("=")
                    saved[key] = os.environ.get(key)
                    os.environ[key] = val
                best[vi] = min(best[vi], timed())
                for key, old in saved.items():
                    if old is None:
                        os.environ.pop(key, None)
                    else:
                        os.environ[key] = old
        vms = best
        ms = best[-1]
    else:
        ms = timed()
    name = "%s %s k%d s%d %dx%d %d>%d b%d%s%s%s" % (op, "deconv" if deconv else "conv", k, s, H, H, Ci, Co, B,
                                                   " reflect" if pad == "reflect" else "", " +res" if residual else "",
                                                   " +mask" if relu_mask else "")
    return {"layer": name, "ms": ms, "tflops": flops / ms / 1e9, "variants_ms": vms}


specs = {"short": SHORT, "deep": DEEP, "all": SHORT + DEEP}[args.set]
if args.only >= 0:
    specs = [specs[args.only]]
out = []
for sp in specs:
    r = run(sp)
    out.append(r)
    print("%-58s %8.3f ms %8.1f TFLOP/s  %s" % (r["layer"], r["ms"], r["tflops"], " ".join("%.3f" % v for v in r["variants_ms"])), flush=True)
print("TOTAL %.3f ms" % sum(r["ms"] for r in out))
if args.json:
    json.dump(out, open(args.json, "w"), indent=1)
