"""GPU parity tests of the entry points added in round 2, each through the C ABI against a plain torch reference of the same
arithmetic: grouped fully-connected layers (csrc/fc_group.cu), channel-slice copy and slice max-pooling (the fused f / g / h
attention projection), and the hoisted conditional-BN projections / folded attention output against the un-fused operator
path they replace."""
import ctypes

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def _items(n):
    from biggan_tensorflow_b200.ops import FcGroupItem
    return (FcGroupItem * n)()


@pytest.mark.parametrize("M", [2, 32, 200])
@pytest.mark.parametrize("wdtype", [torch.bfloat16, torch.float32])
def test_fc_group_forward_and_backward(M, wdtype):
    from biggan_tensorflow_b200 import _lib
    from biggan_tensorflow_b200._lib import call
    from biggan_tensorflow_b200.tensor import stream
    g = torch.Generator(device="cuda").manual_seed(0)
    probs = [(52, 32), (1032, 96), (1032, 1536), (24, 256), (1032, 80)]       # (K, N)
    xs, ws, bs, ys = [], [], [], []
    arr = _items(len(probs))
    for it, (K, N) in zip(arr, probs):
        x = torch.randn(M, K, device="cuda", generator=g)
        wt = (torch.randn(N, K, device="cuda", generator=g) * 0.05).to(wdtype)        # K-major [N][K]
        b = torch.randn(N, device="cuda", generator=g)
        y = torch.empty(M, N, device="cuda")
        it.x, it.w, it.bias, it.y, it.K, it.N, it.ldw = x.data_ptr(), wt.data_ptr(), b.data_ptr(), y.data_ptr(), K, N, K
        xs.append(x); ws.append(wt); bs.append(b); ys.append(y)
    call("bg_fc_group_fwd", arr, len(probs), M, _lib.BG_BF16 if wdtype == torch.bfloat16 else _lib.BG_F32, stream())
    torch.cuda.synchronize()
    for x, wt, b, y in zip(xs, ws, bs, ys):
        ref = x.double() @ wt.double().t() + b.double()
        assert float((y.double() - ref).abs().max() / ref.abs().max()) <= 2e-6
    if wdtype == torch.float32:
        return
    arr2 = _items(len(probs))
    keep = []
    for it, x, (K, N) in zip(arr2, xs, probs):
        dy = torch.randn(M, N, device="cuda", generator=g)
        ld = N + 16                                                                # padded row stride of the destination
        dw = torch.full((K, ld), 7.0, device="cuda")
        db = torch.empty(N, device="cuda")
        it.x, it.w, it.bias, it.y, it.K, it.N, it.ldw = x.data_ptr(), dw.data_ptr(), db.data_ptr(), dy.data_ptr(), K, N, ld
        keep.append((dy, dw, db))
    call("bg_fc_group_bwd", arr2, len(probs), M, stream())
    torch.cuda.synchronize()
    for x, (dy, dw, db), (K, N) in zip(xs, keep, probs):
        ref = x.double().t() @ dy.double()
        assert float((dw[:, :N].double() - ref).abs().max() / ref.abs().max()) <= 2e-6
        assert torch.all(dw[:, N:] == 7.0)                                         # nothing written beyond N
        assert float((db.double() - dy.double().sum(0)).abs().max() / dy.double().sum(0).abs().max()) <= 2e-6


@pytest.mark.parametrize("dtype,width,c0,Nt", [(torch.bfloat16, 16, 16, 80), (torch.bfloat16, 48, 32, 80), (torch.bfloat16, 3, 5, 19),
                                              (torch.float32, 8, 4, 24), (torch.float32, 5, 1, 7)])
def test_copy_channels(dtype, width, c0, Nt):
    from biggan_tensorflow_b200._lib import call
    from biggan_tensorflow_b200.tensor import bg_dtype, stream
    npix = 1000
    src = torch.randn(npix, Nt, device="cuda").to(dtype)
    dst = torch.zeros(npix, width, device="cuda", dtype=dtype)
    call("bg_copy_channels", src.view(-1)[c0:], Nt, dst, width, npix, width, bg_dtype(dtype), stream())
    assert torch.equal(dst, src[:, c0:c0 + width])
    back = torch.full((npix, Nt), 3.0, device="cuda", dtype=dtype)
    call("bg_copy_channels", dst, width, back.view(-1)[c0:], Nt, npix, width, bg_dtype(dtype), stream())
    assert torch.equal(back[:, c0:c0 + width], dst) and torch.all(back[:, :c0] == 3.0) and torch.all(back[:, c0 + width:] == 3.0)


@pytest.mark.parametrize("C,c0,Nt", [(16, 0, 80), (48, 32, 80), (96, 64, 160)])
def test_maxpool2_slice_matches_the_dense_pooling(C, c0, Nt):
    from biggan_tensorflow_b200 import _lib
    from biggan_tensorflow_b200._lib import call
    from biggan_tensorflow_b200.tensor import stream
    B, H, W = 3, 8, 12
    y = torch.randn(B, H, W, Nt, device="cuda").bfloat16()
    dense = y[..., c0:c0 + C].contiguous()
    want = torch.empty(B, H // 2, W // 2, C, device="cuda", dtype=torch.bfloat16)
    call("bg_maxpool2_fwd", dense, want, B, H, W, C, _lib.BG_BF16, stream())
    got = torch.empty_like(want)
    call("bg_maxpool2_slice_fwd", y.view(-1)[c0:], Nt, got, B, H, W, C, _lib.BG_BF16, stream())
    assert torch.equal(got, want)
    dy = torch.randn_like(want)
    dwant = torch.empty_like(dense)
    call("bg_maxpool2_bwd", dense, dy, dwant, B, H, W, C, _lib.BG_BF16, stream())
    dgot = torch.full_like(y, 5.0)
    call("bg_maxpool2_slice_bwd", y.view(-1)[c0:], Nt, dy, dgot.view(-1)[c0:], Nt, B, H, W, C, _lib.BG_BF16, stream())
    assert torch.equal(dgot[..., c0:c0 + C], dwant)
    assert torch.all(dgot[..., :c0] == 5.0) and torch.all(dgot[..., c0 + C:] == 5.0)


def _sa2_grads(monkeypatch, env):
    """self_attention_2 forward + backward in bf16 mode with the given fusion switches; returns output, input gradient
    and every variable gradient."""
    from helpers import dev_tensor, new_product_store, seed_backward
    from biggan_tensorflow_b200 import ops
    from test_ops_gpu import _opt
    for k in ("BG_NO_QKV_FUSION", "BG_NO_GAMMA_FOLD"):
        monkeypatch.delenv(k, raising=False)
    for k in env:
        monkeypatch.setenv(k, "1")
    st = new_product_store(torch.bfloat16)
    st.load({"self_attention/gamma": np.asarray([0.3], np.float32)})
    rng = np.random.default_rng(3)
    x = dev_tensor(rng.standard_normal((2, 16, 16, 64)).astype(np.float32), torch.bfloat16)
    y = ops.self_attention_2(x, 64, _opt()[0])
    seed_backward(y, rng.standard_normal(y.shape).astype(np.float32))
    return y.numpy(), x.grad.float().cpu().numpy(), {k: v.grad.float().cpu().numpy() for k, v in st.vars.items() if v.grad is not None}


def test_fused_attention_projection_and_folded_output_equal_the_unfused_operators(monkeypatch):
    """The stacked f/g/h GEMM + slice pooling and the gamma-folded attn_conv against the operator-by-operator path (three
    convs, max_pooling, conv, gamma * o + x) on the same bf16 kernels: same values up to bf16 rounding of intermediates."""
    ya, gxa, gva = _sa2_grads(monkeypatch, ())
    yb, gxb, gvb = _sa2_grads(monkeypatch, ("BG_NO_QKV_FUSION", "BG_NO_GAMMA_FOLD"))
    rel = lambda a, b: float(np.linalg.norm(a.astype(np.float64) - b) / max(np.linalg.norm(b), 1e-12))
    assert rel(ya, yb) <= 1e-2 and rel(gxa, gxb) <= 2e-2
    assert sorted(gva) == sorted(gvb)
    scale = max(float(np.abs(v).max()) for v in gvb.values())
    for k in gvb:
        if k.endswith("f_conv/bias"):
            continue                                   # analytically zero
        assert np.linalg.norm(gva[k] - gvb[k]) <= 3e-2 * np.linalg.norm(gvb[k]) + 2e-3 * scale, k
