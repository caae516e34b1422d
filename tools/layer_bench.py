"""Single-layer conv micro-benchmark through the C ABI (bf16 tensor-core path): time, TFLOP/s and activation GB/s of
bg_conv2d_fwd / bg_conv2d_dgrad / bg_conv2d_wgrad for one geometry.  Random data; timing only (parity is tests/).
Usage (B200):  python tools/layer_bench.py --op fwd --B 256 --H 128 --Ci 96 --Co 16 --k 3 --stride 1 [--deconv] [--pad zero]
Run under ncu with --iters 1 --warmup 1 to capture the kernels of exactly one layer."""
import argparse
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

ap = argparse.ArgumentParser()
ap.add_argument("--op", default="fwd", choices=["fwd", "dgrad", "wgrad"])
ap.add_argument("--B", type=int, default=256)
ap.add_argument("--H", type=int, default=128)
ap.add_argument("--Ci", type=int, default=96)
ap.add_argument("--Co", type=int, default=96)
ap.add_argument("--k", type=int, default=3)
ap.add_argument("--stride", type=int, default=1)
ap.add_argument("--deconv", action="store_true")
ap.add_argument("--pad", default="reflect", choices=["reflect", "zero"])
ap.add_argument("--residual", action="store_true", help="fused residual add (fwd) / gradient accumulate (dgrad)")
ap.add_argument("--relu-mask", action="store_true", help="dgrad of a reflect-padded conv: fused ReLU mask of the conv's input")
ap.add_argument("--iters", type=int, default=20)
ap.add_argument("--warmup", type=int, default=3)
args = ap.parse_args()

from biggan_tensorflow_b200 import _lib
from biggan_tensorflow_b200._lib import ConvDesc, call, query
from biggan_tensorflow_b200.tensor import stream

dev = torch.device("cuda", 0)
B, H, k, s, Ci, Co = args.B, args.H, args.k, args.stride, args.Ci, args.Co
if args.deconv:
    Ho = H * s
    tot = max((H - 1) * s + k - Ho, 0)
    p = tot // 2
    d = ConvDesc(B, H, H, Ci, Ho, Ho, Co, k, k, s, p, p, tot - p, tot - p, _lib.PAD_ZERO, 1)
    w = torch.randn(k, k, Co, Ci, device=dev).mul_(0.05).bfloat16()
else:
    pad = (k - 1) // 2
    tot = 2 * pad if H % s == 0 else max(k - H % s, 0)
    pt = tot // 2
    Ho = (H + tot - k) // s + 1
    d = ConvDesc(B, H, H, Ci, Ho, Ho, Co, k, k, s, pt, pt, tot - pt, tot - pt,
                 _lib.PAD_REFLECT if (args.pad == "reflect" and tot) else _lib.PAD_ZERO, 0)
    w = torch.randn(k, k, Ci, Co, device=dev).mul_(0.05).bfloat16()
w_t = w.transpose(2, 3).contiguous()
x = torch.randn(B, H, H, Ci, device=dev).bfloat16()
y = torch.empty(B, Ho, Ho, Co, device=dev, dtype=torch.bfloat16)
dy = torch.randn(B, Ho, Ho, Co, device=dev).bfloat16()
dx = torch.empty_like(x)
dw = torch.empty(w.shape, device=dev, dtype=torch.float32)
dt = _lib.BG_BF16
flops = 2.0 * B * (H * H if args.deconv else Ho * Ho) * k * k * Ci * Co


def ws_for(name):
    n = query(name, d, dt)
    return torch.empty(max(int(n), 16), dtype=torch.uint8, device=dev), int(n)


if args.op == "fwd":
    ws, n = ws_for("bg_conv2d_fwd_workspace_bytes")
    res = torch.randn_like(y) if args.residual else None
    fn = lambda: call("bg_conv2d_fwd", d, x, w, w_t, None, res, y, _lib.ACT_NONE, dt, ws if n else None, n, 0, stream())
elif args.op == "dgrad":
    ws, n = ws_for("bg_conv2d_dgrad_workspace_bytes")
    res = torch.randn_like(dx) if args.residual else None
    mask = torch.randn_like(dx) if args.relu_mask else None
    fn = lambda: call("bg_conv2d_dgrad", d, dy, w, w_t, res, mask, dx, dt, ws if n else None, n, 0, stream())
else:
    ws, n = ws_for("bg_conv2d_wgrad_workspace_bytes")
    fn = lambda: call("bg_conv2d_wgrad", d, x, dy, dw, dt, ws, ws.numel(), 0, stream())

for _ in range(args.warmup):
    fn()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(args.iters):
    fn()
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / args.iters
act_bytes = 2.0 * (x.numel() + y.numel())
import ctypes
c = (ctypes.c_int * 2)()
_lib.lib.bg_debug_counters(ctypes.addressof(c), ctypes.addressof(c) + 4)
print("%s %s k%d s%d %dx%d %d>%d b%d: %.3f ms  %.1f TFLOP/s  act %.0f GB/s  (tc launches fwd=%d wgrad=%d)" %
      (args.op, "deconv" if args.deconv else "conv", k, s, H, H, Ci, Co, B, ms, flops / ms / 1e9, act_bytes / ms / 1e6, c[0], c[1]))
