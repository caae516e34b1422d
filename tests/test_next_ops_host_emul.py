"""CPU execution check of the SURVEY 8(f) kernels (`csrc/next_ops.cu`): the SAME source is compiled for the host with
`-DBG_HOST_EMULATION` (tests/host_emul/cuda_host_emul.h: blocks run in turn, the threads of a block in turn or -- when the kernel synchronises -- as round-robin fibers, so
`__syncthreads` / warp shuffles behave) and called through the same C entry points, argument for argument, against the
oracle.  This was written when no GPU time was left in round 1: it validates index arithmetic, launch geometry and the
fixed-order reductions of the real kernel code; the GPU parity tests (tests/test_zz_reference_golden_gpu.py, gated) still
have to run on hardware."""
import ctypes
import os
import subprocess

import numpy as np
import pytest
import torch

from oracle import ref_ops as R

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
F32, BF16 = 0, 1


@pytest.fixture(scope="module")
def emul(tmp_path_factory):
    from biggan_tensorflow_b200._lib import parse_header
    so = str(tmp_path_factory.mktemp("host_emul") / "next_ops_host.so")
    src = os.path.join(ROOT, "biggan_tensorflow_b200", "csrc", "next_ops.cu")
    subprocess.run(["g++", "-std=c++20", "-O1", "-pthread", "-DBG_HOST_EMULATION", "-x", "c++", "-shared", "-fPIC", src, "-o", so],
                   check=True)
    lib = ctypes.CDLL(so)
    for name, (res, args) in parse_header().items():
        if hasattr(lib, name):
            fn = getattr(lib, name)
            fn.restype, fn.argtypes = res, args
    return lib


def to_dev(a, dtype):
    """numpy float32 -> buffer in the storage dtype (bf16 as uint16 bit patterns)."""
    a = np.ascontiguousarray(a, dtype=np.float32)
    if dtype == F32:
        return a.copy()
    return torch.from_numpy(a).bfloat16().view(torch.int16).numpy().copy()


def from_dev(buf, dtype):
    if dtype == F32:
        return buf.astype(np.float64)
    return torch.from_numpy(buf).view(torch.bfloat16).double().numpy()


def ptr(a):
    return None if a is None else a.ctypes.data_as(ctypes.c_void_p)


def close(a, b, dtype, what):
    tol = 1e-6 if dtype == F32 else 1.2e-2
    scale = max(float(np.abs(b).max()), 1e-6)
    assert float(np.abs(np.asarray(a, np.float64) - b).max()) <= tol * scale, what


DTYPES = [F32, BF16]


@pytest.mark.parametrize("dtype", DTYPES)
def test_lrelu(emul, dtype):
    x = np.random.default_rng(0).standard_normal(3001).astype(np.float32)
    dy = np.random.default_rng(1).standard_normal(3001).astype(np.float32)
    xd, dyd = to_dev(x, dtype), to_dev(dy, dtype)
    y, dx = np.zeros_like(xd), np.zeros_like(xd)
    assert emul.bg_lrelu_fwd(ptr(xd), ptr(y), x.size, 0.2, dtype, None) == 0
    assert emul.bg_lrelu_bwd(ptr(xd), ptr(dyd), ptr(dx), x.size, 0.2, dtype, None) == 0
    xt = torch.tensor(from_dev(xd, dtype), requires_grad=True)
    yr = R.lrelu(xt, 0.2)
    (g,) = torch.autograd.grad((yr * torch.tensor(from_dev(dyd, dtype))).sum(), [xt])
    close(from_dev(y, dtype), yr.detach().numpy(), dtype, "lrelu fwd")
    close(from_dev(dx, dtype), g.numpy(), dtype, "lrelu bwd")


@pytest.mark.parametrize("dtype", DTYPES)
@pytest.mark.parametrize("shape", [(2, 5, 7, 3), (3, 4, 4, 40), (1, 1, 1, 300)])
def test_prelu(emul, dtype, shape):
    rng = np.random.default_rng(2)
    C = shape[-1]
    npix = int(np.prod(shape[:-1]))
    x, dy = rng.standard_normal(shape).astype(np.float32), rng.standard_normal(shape).astype(np.float32)
    alpha = rng.standard_normal(C).astype(np.float32)
    xd, dyd = to_dev(x, dtype), to_dev(dy, dtype)
    y, dx = np.zeros_like(xd), np.zeros_like(xd)
    da = np.full(C, 7.0, np.float32)
    ws = np.zeros(emul.bg_prelu_workspace_bytes(npix, C) // 4 + 1, np.float32)
    assert emul.bg_prelu_fwd(ptr(xd), ptr(alpha), ptr(y), npix, C, dtype, None) == 0
    assert emul.bg_prelu_bwd(ptr(xd), ptr(alpha), ptr(dyd), ptr(dx), ptr(da), 1, npix, C, dtype, ptr(ws), ws.nbytes, None) == 0
    st = R.RefStore(torch.float64, params={"prelu/alpha": alpha})
    xt = torch.tensor(from_dev(xd, dtype), requires_grad=True)
    yr = R.make_prelu(st)(xt)
    g, ga = torch.autograd.grad((yr * torch.tensor(from_dev(dyd, dtype))).sum(), [xt, st.vars["prelu/alpha"]])
    close(from_dev(y, dtype), yr.detach().numpy(), dtype, "prelu fwd")
    close(from_dev(dx, dtype), g.numpy(), dtype, "prelu dx")
    close(da.astype(np.float64) - 7.0, ga.numpy(), F32 if dtype == F32 else BF16, "prelu dalpha (accumulated onto 7)")
    da2 = np.full(C, 7.0, np.float32)                      # accumulate = 0 overwrites; dx = NULL is allowed
    assert emul.bg_prelu_bwd(ptr(xd), ptr(alpha), ptr(dyd), None, ptr(da2), 0, npix, C, dtype, ptr(ws), ws.nbytes, None) == 0
    close(da2, ga.numpy(), F32 if dtype == F32 else BF16, "prelu dalpha")
    assert emul.bg_prelu_bwd(ptr(xd), ptr(alpha), ptr(dyd), None, ptr(da2), 0, npix, C, dtype, ptr(ws), 4, None) != 0   # workspace check


@pytest.mark.parametrize("dtype", DTYPES)
@pytest.mark.parametrize("shape", [(2, 4, 6, 5), (1, 8, 2, 33)])
def test_avg_pool_and_up_sample(emul, dtype, shape):
    rng = np.random.default_rng(3)
    B, H, W, C = shape
    x = rng.standard_normal(shape).astype(np.float32)
    xd = to_dev(x, dtype)
    xt = torch.tensor(from_dev(xd, dtype), requires_grad=True)
    # average pooling and its backward
    y = np.zeros((B, H // 2, W // 2, C), xd.dtype)
    assert emul.bg_avgpool2_fwd(ptr(xd), ptr(y), B, H, W, C, dtype, None) == 0
    yr = R.avg_pooling(xt)
    close(from_dev(y, dtype), yr.detach().numpy(), dtype, "avgpool fwd")
    dy = rng.standard_normal(y.shape).astype(np.float32)
    dyd, dx = to_dev(dy, dtype), np.zeros_like(xd)
    assert emul.bg_avgpool2_bwd(ptr(dyd), ptr(dx), B, H, W, C, dtype, None) == 0
    (g,) = torch.autograd.grad((yr * torch.tensor(from_dev(dyd, dtype))).sum(), [xt])
    close(from_dev(dx, dtype), g.numpy(), dtype, "avgpool bwd")
    assert emul.bg_avgpool2_fwd(ptr(xd), ptr(y), B, H + 1, W, C, dtype, None) != 0           # odd maps are refused
    # nearest-neighbour up-sampling and its backward
    u = np.zeros((B, 2 * H, 2 * W, C), xd.dtype)
    assert emul.bg_upsample2_fwd(ptr(xd), ptr(u), B, H, W, C, dtype, None) == 0
    ur = R.up_sample(xt, 2)
    close(from_dev(u, dtype), ur.detach().numpy(), dtype, "upsample fwd")
    du = rng.standard_normal(u.shape).astype(np.float32)
    dud, dx2 = to_dev(du, dtype), np.zeros_like(xd)
    assert emul.bg_upsample2_bwd(ptr(dud), ptr(dx2), B, H, W, C, dtype, None) == 0
    (g2,) = torch.autograd.grad((ur * torch.tensor(from_dev(dud, dtype))).sum(), [xt])
    close(from_dev(dx2, dtype), g2.numpy(), dtype, "upsample bwd")


@pytest.mark.parametrize("dtype", DTYPES)
@pytest.mark.parametrize("shape,r", [((2, 3, 4, 8), 2), ((1, 2, 2, 27), 3), ((2, 2, 3, 4), 1)])
def test_depth_to_space_and_its_adjoint(emul, dtype, shape, r):
    rng = np.random.default_rng(4)
    B, H, W, C = shape
    x = rng.standard_normal(shape).astype(np.float32)
    xd = to_dev(x, dtype)
    y = np.zeros((B, H * r, W * r, C // (r * r)), xd.dtype)
    assert emul.bg_depth_to_space(ptr(xd), ptr(y), B, H, W, C, r, 0, dtype, None) == 0
    yr = R.depth_to_space(torch.tensor(from_dev(xd, dtype)), r).numpy()
    assert np.array_equal(from_dev(y, dtype), yr)                                            # a pure permutation: exact
    back = np.zeros_like(xd)
    assert emul.bg_depth_to_space(ptr(y), ptr(back), B, H, W, C, r, 1, dtype, None) == 0
    assert np.array_equal(back, xd)
    if r > 1:
        assert emul.bg_depth_to_space(ptr(xd), ptr(y), B, H, W, C + 1, r, 0, dtype, None) != 0


@pytest.mark.parametrize("kind,K,C", [("ortho", 40, 12), ("ortho", 30, 70), ("ortho", 5, 300),          # C > 256: strided row loops
                                      ("ortho_cosine", 40, 12), ("ortho_cosine", 30, 70), ("ortho_cosine", 3, 260)])
def test_orthogonal_regularizer_rows(emul, kind, K, C):
    """bg_ortho_reg_rows on G = W^T W: loss and, through dW = W (dG + dG^T), the gradient of utils.py:185-235."""
    rng = np.random.default_rng(5)
    W = (rng.standard_normal((K, C)) * 0.3).astype(np.float32)
    G = (W.astype(np.float64).T @ W.astype(np.float64)).astype(np.float32)
    dG, rowloss = np.zeros((C, C), np.float32), np.zeros(C, np.float32)
    loss = np.array([2.5], np.float32)
    scale = 1e-2
    assert emul.bg_ortho_reg_rows(ptr(G), ptr(dG), ptr(rowloss), ptr(loss), 1, C, scale, {"ortho": 0, "ortho_cosine": 1}[kind], None) == 0
    wt = torch.tensor(W.astype(np.float64), requires_grad=True)
    L = R.orthogonal_regularizer(scale, type=kind)(wt)
    (gw,) = torch.autograd.grad(L, [wt])
    assert abs(float(loss[0]) - 2.5 - float(L.detach())) <= 2e-5 * max(float(L.detach()), 1e-3) + 1e-6
    dW = W.astype(np.float64) @ (dG.astype(np.float64) + dG.astype(np.float64).T)
    assert np.abs(dW - gw.numpy()).max() <= 2e-4 * np.abs(gw.numpy()).max()
    assert emul.bg_ortho_reg_rows(ptr(G), ptr(dG), ptr(rowloss), ptr(loss), 0, C, scale, 2, None) != 0   # "Unknown regularization method."


@pytest.mark.parametrize("shared", [False, True])
def test_renorm_factors_and_affine(emul, shared):
    """ops.py:684-707: r, d, running-statistics / fade-in updates, gamma' = r gamma, beta' = beta + d gamma and its adjoint."""
    rng = np.random.default_rng(6)
    C, B = 150, 5
    mean, var = rng.standard_normal(C).astype(np.float32), (rng.random(C) + 0.2).astype(np.float32)
    rmean, rvar = rng.standard_normal(C).astype(np.float32), (rng.random(C) * 3 + 0.05).astype(np.float32)
    w = np.array([0.7], np.float32)
    rm0, rv0 = rmean.copy(), rvar.copy()
    r, d, w_out = np.zeros(C, np.float32), np.zeros(C, np.float32), np.zeros(1, np.float32)
    eps, rmin, rmax, dmax, decay, fade = 1e-5, 1 / 1.2, 1.2, 0.3, 0.8, 0.99
    rc = emul.bg_renorm_factors(ptr(mean), ptr(var), ptr(rmean), ptr(rvar), None if shared else ptr(w), None if shared else ptr(w_out),
                                ptr(r), ptr(d), C, eps, rmin, rmax, dmax, decay, 1 - decay, fade, 1 - fade, 0 if shared else 1, None)
    assert rc == 0
    ww = 1.0 if shared else 0.7
    sigma, rsigma = np.sqrt(var.astype(np.float64) + eps), np.sqrt(rv0.astype(np.float64) + eps)
    wsig, wmean = ww * rsigma + (1 - ww) * sigma, ww * rm0 + (1 - ww) * mean
    assert np.allclose(r, np.clip(sigma / wsig, rmin, rmax), rtol=1e-5) and np.allclose(d, np.clip((mean - wmean) / wsig, -dmax, dmax), rtol=1e-4, atol=1e-6)
    assert (r == np.float32(rmax)).any() and (d == np.float32(-dmax)).any()              # the clips are exercised
    if shared:
        assert np.array_equal(rmean, rm0) and np.array_equal(rvar, rv0)
    else:
        assert np.allclose(rmean, rm0 * decay + mean * (1 - decay), rtol=1e-6, atol=1e-7) and np.allclose(rvar, rv0 * decay + var * (1 - decay), rtol=1e-6)
        assert abs(float(w_out[0]) - (0.7 * fade + 1 - fade)) < 1e-6
    gamma, beta = rng.standard_normal((B, C)).astype(np.float32), rng.standard_normal((B, C)).astype(np.float32)
    g2, b2 = np.zeros_like(gamma), np.zeros_like(beta)
    assert emul.bg_renorm_affine_fwd(ptr(gamma), ptr(beta), ptr(r), ptr(d), ptr(g2), ptr(b2), B, C, None) == 0
    assert np.allclose(g2, r * gamma, rtol=1e-6) and np.allclose(b2, beta + d * gamma, rtol=1e-6, atol=1e-6)
    dg2, db2 = rng.standard_normal((B, C)).astype(np.float32), rng.standard_normal((B, C)).astype(np.float32)
    dg = np.zeros_like(gamma)
    assert emul.bg_renorm_affine_bwd(ptr(dg2), ptr(db2), ptr(r), ptr(d), ptr(dg), B, C, None) == 0
    assert np.allclose(dg, r * dg2 + d * db2, rtol=1e-5, atol=1e-6)
