"""bg_conv2d_dgrad / bg_conv2d_fwd through the C ABI at shapes whose tiles exercise every epilogue path of the tcgen05 forward-type
kernel (csrc/conv_tc.cu): fold mode (reflect-padded dgrad, ops.py:65-84) with tiles made of interior pixels only (bulk tensor
stores), mixed tiles (direct stores + pad ring), fused gradient accumulation and the ReLU mask of the conv's input, all residual /
mask vectors of a half-tile requested up front (BN <= 96) or in rolling groups (BN > 96), the mask on an unpadded 1x1 dgrad, and the
per-tap folded dy path of the small maps (H <= 4).
Checker: torch float32 autograd of the same convolution on the same bf16-rounded operands (test infrastructure only)."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def _desc(B, H, Ci, Co, k, s, reflect):
    from biggan_tensorflow_b200 import _lib
    pd = (k - 1) // 2
    tot = 2 * pd if H % s == 0 else max(k - H % s, 0)
    pt = tot // 2
    Ho = (H + tot - k) // s + 1
    return _lib.ConvDesc(B, H, H, Ci, Ho, Ho, Co, k, k, s, pt, pt, tot - pt, tot - pt,
                         _lib.PAD_REFLECT if (reflect and tot) else _lib.PAD_ZERO, 0), Ho, (pt, tot - pt)


def _ref_conv(x, w, s, pads, reflect):
    """x [B,H,W,Ci], w HWIO (float32 cuda) -> y [B,Ho,Wo,Co]"""
    xp = x.permute(0, 3, 1, 2)
    if pads[0] or pads[1]:
        xp = torch.nn.functional.pad(xp, (pads[0], pads[1], pads[0], pads[1]), mode="reflect" if reflect else "constant")
    return torch.nn.functional.conv2d(xp, w.permute(3, 2, 0, 1), stride=s).permute(0, 2, 3, 1)


DGRAD_CASES = [
    # B, H, Ci, Co, k, s, reflect
    (32, 16, 96, 96, 3, 1, True),       # padded 18x18 -> box (2,2,32): interior tiles (bulk stores) and ring tiles (direct) in one launch
    (32, 16, 96, 192, 3, 2, True),      # strided: 4 sub-pixel phases, each with its own interior offset
    (8, 64, 96, 96, 3, 1, True),        # the D block-1 geometry at a small batch
    (32, 8, 192, 192, 3, 1, True),      # BN = 192: rolling residual groups, mask inline
    (4, 64, 96, 32, 1, 1, False),       # unpadded 1x1: mask in the plain epilogue (tap-stacked RGB output conv)
    (3, 33, 48, 80, 3, 1, True),        # ragged: odd size, BN = 48
    (64, 4, 256, 256, 3, 1, True),      # 4x4 map: dy folded per tap (reflect_dyfold_kernel), 9-tap GEMM on the unpadded domain
    (24, 4, 512, 768, 3, 1, True),      # ... three n-tiles of 256, batch not a multiple of the 8-image box
    (7, 3, 64, 96, 3, 1, True),         # 3x3 map: pixel 1 is the mirror image of both pad rows
    (5, 2, 32, 64, 3, 1, True),         # 2x2 map
]


@pytest.mark.parametrize("fold_store", ["1", "0"])
@pytest.mark.parametrize("res,mask", [(0, 0), (0, 1), (1, 0), (1, 1)])
@pytest.mark.parametrize("case", DGRAD_CASES)
def test_dgrad_epilogue_paths(case, res, mask, fold_store, monkeypatch):
    from biggan_tensorflow_b200 import _lib
    from biggan_tensorflow_b200._lib import call, query
    from biggan_tensorflow_b200.tensor import stream
    B, H, Ci, Co, k, s, reflect = case
    if fold_store == "0" and not reflect:
        pytest.skip("knob only affects the fold mode")
    monkeypatch.setenv("BG_TC_TMA_STORE_FOLD", fold_store)
    dev = torch.device("cuda", 0)
    d, Ho, pads = _desc(B, H, Ci, Co, k, s, reflect)
    g = torch.Generator(device="cpu").manual_seed(1234 + B + H + Ci)
    w = (torch.randn(k, k, Ci, Co, generator=g) * 0.05).bfloat16().to(dev)
    x = torch.randn(B, H, H, Ci, generator=g).bfloat16().to(dev)
    dy = torch.randn(B, Ho, Ho, Co, generator=g).bfloat16().to(dev)
    r = torch.randn(B, H, H, Ci, generator=g).bfloat16().to(dev) if res else None
    m = x if mask else None
    w_t = w.transpose(2, 3).contiguous()
    dx = torch.full_like(x, float("nan"))
    n = int(query("bg_conv2d_dgrad_workspace_bytes", d, _lib.BG_BF16))
    ws = torch.empty(max(n, 16), dtype=torch.uint8, device=dev)
    call("bg_conv2d_dgrad", d, dy, w, w_t, r, m, dx, _lib.BG_BF16, ws if n else None, n, 0, stream())
    torch.cuda.synchronize()
    xr = x.float().requires_grad_(True)
    (_ref_conv(xr, w.float(), s, pads, reflect) * dy.float()).sum().backward()
    want = xr.grad
    if mask:
        want = want * (x.float() > 0)
    if res:
        want = want + r.float()
    got = dx.float()
    assert bool(torch.isfinite(got).all())
    err = float((got - want).norm() / want.norm().clamp_min(1e-6))
    worst = float((got - want).abs().max() / want.abs().max())
    assert err <= 6e-3 and worst <= 2e-2, (case, res, mask, err, worst)


@pytest.mark.parametrize("case", [(16, 32, 96, 96, 3, 1, True), (8, 64, 48, 96, 1, 1, False), (4, 64, 192, 160, 1, 1, False)])
@pytest.mark.parametrize("res", [0, 1])
def test_fwd_epilogue_paths(case, res):
    """forward with bias + fused residual + ReLU: batched (BN <= 96) and rolling (BN > 96) residual loads, bulk-store epilogue"""
    from biggan_tensorflow_b200 import _lib
    from biggan_tensorflow_b200._lib import call, query
    from biggan_tensorflow_b200.tensor import stream
    B, H, Ci, Co, k, s, reflect = case
    dev = torch.device("cuda", 0)
    d, Ho, pads = _desc(B, H, Ci, Co, k, s, reflect)
    g = torch.Generator(device="cpu").manual_seed(99 + Ci)
    w = (torch.randn(k, k, Ci, Co, generator=g) * 0.05).bfloat16().to(dev)
    x = torch.randn(B, H, H, Ci, generator=g).bfloat16().to(dev)
    bias = torch.randn(Co, generator=g).to(dev)
    r = torch.randn(B, Ho, Ho, Co, generator=g).bfloat16().to(dev) if res else None
    w_t = w.transpose(2, 3).contiguous()
    y = torch.full((B, Ho, Ho, Co), float("nan"), dtype=torch.bfloat16, device=dev)
    n = int(query("bg_conv2d_fwd_workspace_bytes", d, _lib.BG_BF16))
    ws = torch.empty(max(n, 16), dtype=torch.uint8, device=dev)
    call("bg_conv2d_fwd", d, x, w, w_t, bias, r, y, _lib.ACT_RELU, _lib.BG_BF16, ws if n else None, n, 0, stream())
    torch.cuda.synchronize()
    want = _ref_conv(x.float(), w.float(), s, pads, reflect) + bias
    if res:
        want = want + r.float()
    want = torch.relu(want)
    got = y.float()
    assert bool(torch.isfinite(got).all())
    err = float((got - want).norm() / want.norm().clamp_min(1e-6))
    assert err <= 6e-3, (case, res, err)
