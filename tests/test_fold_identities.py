"""CPU statement of the identity behind the small-map dgrad path (csrc/conv_simt.cu: reflect_dyfold*_kernel):
for a reflect-padded 3x3 / stride 1 / pad 1 conv (ops.py:65-84) the fold of the padded-domain input gradient commutes with the
per-pixel weight contraction,
    dx[i] = sum_t dyf_t[i] W_t^T,   dyf_t[i] = sum over padded p that reflect onto i of dy[p - tap_t]   (zero outside the map),
so dx is a 9-"tap" GEMM over the UNPADDED domain whose tap t reads its own pre-folded copy of dy.  Checked against torch autograd of
the reflect-padded convolution in float64 (the kernels themselves are checked on the GPU: tests/test_conv_epilogue_gpu.py)."""
import numpy as np
import pytest
import torch


def fold_dy(dy):
    """dy [B,H,W,C] -> dyf [9,B,H,W,C], written the way the kernel enumerates it"""
    B, H, W, C = dy.shape
    out = torch.zeros(9, B, H, W, C, dtype=dy.dtype)
    for t in range(9):
        oy, ox = t // 3 - 1, t % 3 - 1
        for i in range(H):
            pys = [i] + ([-1] if i == 1 else []) + ([H] if i == H - 2 else [])
            for j in range(W):
                pxs = [j] + ([-1] if j == 1 else []) + ([W] if j == W - 2 else [])
                for py in pys:
                    for px in pxs:
                        sy, sx = py - oy, px - ox
                        if 0 <= sy < H and 0 <= sx < W:
                            out[t, :, i, j] += dy[:, sy, sx]
    return out


@pytest.mark.parametrize("H,W", [(4, 4), (2, 2), (3, 3), (5, 4), (2, 5), (8, 8)])
def test_dgrad_of_reflect_padded_conv_is_a_gemm_over_per_tap_folded_dy(H, W):
    g = torch.Generator().manual_seed(H * 10 + W)
    B, Ci, Co = 3, 5, 7
    x = torch.randn(B, H, W, Ci, generator=g, dtype=torch.float64, requires_grad=True)
    w = torch.randn(3, 3, Ci, Co, generator=g, dtype=torch.float64)          # HWIO
    dy = torch.randn(B, H, W, Co, generator=g, dtype=torch.float64)
    xp = torch.nn.functional.pad(x.permute(0, 3, 1, 2), (1, 1, 1, 1), mode="reflect")
    y = torch.nn.functional.conv2d(xp, w.permute(3, 2, 0, 1)).permute(0, 2, 3, 1)
    (y * dy).sum().backward()
    dyf = fold_dy(dy)
    dx = sum(torch.einsum("bhwo,io->bhwi", dyf[t], w[t // 3, t % 3]) for t in range(9))
    assert float((dx - x.grad).abs().max()) < 1e-12 * max(1.0, float(x.grad.abs().max()))


def test_the_same_identity_gives_the_weight_gradient_of_the_tap_stacked_rgb_conv():
    """rgb_out.cu: dWall[c, (t, n)] = sum_i x[i, c] dYf[i, (t, n)] on the unpadded domain equals the weight gradient of the padded conv."""
    g = torch.Generator().manual_seed(3)
    B, H, W, Ci = 2, 6, 5, 4
    x = torch.randn(B, H, W, Ci, generator=g, dtype=torch.float64)
    w = torch.randn(3, 3, Ci, 3, generator=g, dtype=torch.float64, requires_grad=True)
    dy = torch.randn(B, H, W, 3, generator=g, dtype=torch.float64)
    xp = torch.nn.functional.pad(x.permute(0, 3, 1, 2), (1, 1, 1, 1), mode="reflect")
    y = torch.nn.functional.conv2d(xp, w.permute(3, 2, 0, 1)).permute(0, 2, 3, 1)
    (y * dy).sum().backward()
    dyf = fold_dy(dy)
    dw = torch.stack([torch.einsum("bhwi,bhwo->io", x, dyf[t]) for t in range(9)]).reshape(3, 3, Ci, 3)
    assert float((dw - w.grad).abs().max()) < 1e-12 * max(1.0, float(w.grad.abs().max()))
