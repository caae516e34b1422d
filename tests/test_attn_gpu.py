"""Fused flash-style attention core (tcgen05) of self_attention_2 (ops.py:481-485) vs a float64 reference."""
import numpy as np
import pytest
import torch

from helpers import assert_close

pytestmark = pytest.mark.gpu


def _ref(q, k, v):
    s = torch.einsum("bic,bjc->bij", q, k)
    p = torch.softmax(s, dim=-1)
    return p @ v, torch.logsumexp(s, dim=-1)


@pytest.mark.parametrize("B,Nq,Nk,dq,dv", [(2, 256, 64, 16, 16), (3, 1024, 256, 32, 96), (2, 4096, 1024, 32, 96),
                                           (2, 4096, 1024, 16, 48), (1, 64, 16, 16, 32), (2, 320, 80, 32, 64)])
def test_attn2_forward(B, Nq, Nk, dq, dv):
    from biggan_tensorflow_b200._lib import call
    from biggan_tensorflow_b200.tensor import stream
    rng = np.random.default_rng(0)
    q = torch.from_numpy(rng.standard_normal((B, Nq, dq)).astype(np.float32)).cuda().bfloat16()
    k = torch.from_numpy(rng.standard_normal((B, Nk, dq)).astype(np.float32)).cuda().bfloat16()
    v = torch.from_numpy(rng.standard_normal((B, Nk, dv)).astype(np.float32)).cuda().bfloat16()
    o = torch.empty(B, Nq, dv, dtype=torch.bfloat16, device="cuda")
    lse = torch.empty(B, Nq, dtype=torch.float32, device="cuda")
    call("bg_attn2_fwd", q, k, v, o, lse, B, Nq, Nk, dq, dv, stream())
    torch.cuda.synchronize()
    ro, rl = _ref(q.double().cpu(), k.double().cpu(), v.double().cpu())
    assert_close(lse.cpu().numpy(), rl.numpy(), 1e-5, "lse")
    assert_close(o.float().cpu().numpy(), ro.numpy(), 2e-2, "o")


@pytest.mark.parametrize("scales", [(0.3, 0.3), (1.0, 1.0), (0.3, 1.5), (0.0, 0.3)], ids=["single-pass", "two-pass", "mixed-batch", "zero-q"])
@pytest.mark.parametrize("B,Nq,Nk,dq,dv", [(2, 4096, 1024, 32, 96), (2, 4096, 1024, 16, 48), (2, 320, 80, 32, 64), (4, 256, 64, 16, 16)])
def test_attn2_forward_with_score_bound(B, Nq, Nk, dq, dv, scales):
    """bg_attn2_fwd_ws: a pre-kernel bounds |q_i . k_j| per image by max|q_i| max|k_j|; images whose bound is <= 30 skip the
    row-maximum pass and shift by the bound instead (same o, same lse), the others take the two-pass path -- decided per image."""
    from biggan_tensorflow_b200._lib import call, query
    from biggan_tensorflow_b200.tensor import stream
    rng = np.random.default_rng(7)
    sc = torch.tensor([scales[i % 2] for i in range(B)], dtype=torch.float32).view(B, 1, 1)
    q = (torch.from_numpy(rng.standard_normal((B, Nq, dq)).astype(np.float32)) * sc).cuda().bfloat16()
    k = (torch.from_numpy(rng.standard_normal((B, Nk, dq)).astype(np.float32)) * sc.clamp_min(0.3)).cuda().bfloat16()
    v = torch.from_numpy(rng.standard_normal((B, Nk, dv)).astype(np.float32)).cuda().bfloat16()
    o = torch.full((B, Nq, dv), float("nan"), dtype=torch.bfloat16, device="cuda")
    lse = torch.full((B, Nq), float("nan"), dtype=torch.float32, device="cuda")
    nb = int(query("bg_attn2_fwd_workspace_bytes", B))
    ws = torch.empty(nb, dtype=torch.uint8, device="cuda")
    call("bg_attn2_fwd_ws", q, k, v, o, lse, B, Nq, Nk, dq, dv, ws, nb, stream())
    torch.cuda.synchronize()
    bound = ws[:4 * B].view(torch.float32).cpu().numpy()
    qn, kn = q.float().norm(dim=-1).amax(dim=1).cpu().numpy(), k.float().norm(dim=-1).amax(dim=1).cpu().numpy()
    assert np.all(bound >= qn * kn) and np.all(bound <= qn * kn * 1.01 + 2e-3)
    if scales == (0.3, 1.5):
        assert bound[0] <= 30.0 < bound[1]              # both paths in one launch
    ro, rl = _ref(q.double().cpu(), k.double().cpu(), v.double().cpu())
    assert_close(lse.cpu().numpy(), rl.numpy(), 1e-5, "lse")
    assert_close(o.float().cpu().numpy(), ro.numpy(), 2e-2, "o")


@pytest.mark.parametrize("fused_dq", [True, False], ids=["dq-in-dkdv", "two-kernels"])
@pytest.mark.parametrize("B,Nq,Nk,dq,dv", [(2, 256, 64, 16, 16), (2, 1024, 256, 32, 96), (1, 4096, 1024, 32, 96),
                                           (2, 4096, 1024, 16, 48), (1, 64, 16, 16, 32), (2, 320, 80, 32, 64), (1, 512, 128, 64, 128)])
def test_attn2_backward(B, Nq, Nk, dq, dv, fused_dq):
    """fused_dq: with the workspace the dK/dV kernel also forms dQ (per-key-block partials + fixed-order reduce); without
    it the separate dQ kernel runs.  Both against the float64 reference."""
    from biggan_tensorflow_b200._lib import call, query
    from biggan_tensorflow_b200.tensor import stream
    rng = np.random.default_rng(1)
    mk = lambda *sh: torch.from_numpy(rng.standard_normal(sh).astype(np.float32)).cuda().bfloat16()
    q, k, v, do = mk(B, Nq, dq), mk(B, Nk, dq), mk(B, Nk, dv), mk(B, Nq, dv)
    o = torch.empty(B, Nq, dv, dtype=torch.bfloat16, device="cuda")
    lse = torch.empty(B, Nq, dtype=torch.float32, device="cuda")
    call("bg_attn2_fwd", q, k, v, o, lse, B, Nq, Nk, dq, dv, stream())
    gq, gk, gv = torch.empty_like(q), torch.empty_like(k), torch.empty_like(v)
    delta = torch.empty(B, Nq, dtype=torch.float32, device="cuda")
    ws = torch.empty(int(query("bg_attn2_bwd_workspace_bytes", B, Nq, Nk, dq)), dtype=torch.uint8, device="cuda") if fused_dq else None
    call("bg_attn2_bwd", q, k, v, o, do, lse, delta, gq, gk, gv, B, Nq, Nk, dq, dv, ws, ws.numel() if fused_dq else 0, stream())
    torch.cuda.synchronize()
    rq, rk, rv = [t.double().cpu().requires_grad_(True) for t in (q, k, v)]
    ro, _ = _ref(rq, rk, rv)
    g = torch.autograd.grad((ro * do.double().cpu()).sum(), [rq, rk, rv])
    for got, ref, name in zip((gq, gk, gv), g, ("dq", "dk", "dv")):
        assert_close(got.float().cpu().numpy(), ref.numpy(), 2e-2, name)
