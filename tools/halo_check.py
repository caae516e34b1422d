"""A/B check of the halo forward kernel against the plain tensor-core kernel and torch on the same data (B200).
Round-1 finding recorded here: the UMMA SWIZZLE_128B pattern is a function of the absolute shared-memory address, so a
descriptor that starts j pixel rows (j * 128 B) into a tile needs NO matrix-base-offset (setting it to j gave garbage)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from biggan_tensorflow_b200 import _lib
from biggan_tensorflow_b200._lib import ConvDesc, call
from biggan_tensorflow_b200.tensor import stream

dev = torch.device("cuda", 0)
torch.manual_seed(0)
for (B, H, Ci, Co) in ((2, 128, 64, 16), (2, 128, 96, 96), (1, 256, 16, 32)):
    d = ConvDesc(B, H, H, Ci, H, H, Co, 3, 3, 1, 1, 1, 1, 1, _lib.PAD_ZERO, 0)
    w = torch.randn(3, 3, Ci, Co, device=dev).mul_(0.05).bfloat16()
    w_t = w.transpose(2, 3).contiguous()
    x = torch.randn(B, H, H, Ci, device=dev).bfloat16()
    outs = {}
    for name, env in (("plain", {"BG_TC_HALO": "0"}), ("halo", {"BG_TC_HALO": "1"})):
        os.environ.update(env)
        y = torch.zeros(B, H, H, Co, device=dev, dtype=torch.bfloat16)
        call("bg_conv2d_fwd", d, x, w, w_t, None, None, y, _lib.ACT_NONE, _lib.BG_BF16, None, 0, 0, stream())
        torch.cuda.synchronize()
        outs[name] = y.float()
    ref = torch.nn.functional.conv2d(x.float().permute(0, 3, 1, 2), w.float().permute(3, 2, 0, 1), padding=1).permute(0, 2, 3, 1)
    for name, y in outs.items():
        print("%dx%d %d>%d %-9s max|y-ref| = %.4f  (ref max %.3f)" % (H, H, Ci, Co, name, float((y - ref).abs().max()), float(ref.abs().max())))
