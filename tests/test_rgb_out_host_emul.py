"""CPU execution check of the tap-stacked RGB output conv kernels (`csrc/rgb_out.cu`, G_logit of BigGAN.py:570 in tensor-core mode):
the SAME kernel source is compiled for the host with `-DBG_HOST_EMULATION` (tests/host_emul/cuda_host_emul.h; warps as fibers, so
the lane shuffles of the tap sum behave) and called through the C entry points against a torch float64 statement of the tap sum
and its autograd adjoint.  Checks the index arithmetic of the real kernels (30-pixel warp segments, 32-byte windows, reflect /
zero borders, the pre-image sets of the adjoint); the GPU parity test is tests/test_ops_gpu.py::test_conv_rgb_out_tap_stacked."""
import ctypes
import os
import subprocess

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ACT_NONE, ACT_RELU, ACT_TANH = 0, 1, 2
PAD_ZERO, PAD_REFLECT = 0, 1


@pytest.fixture(scope="module")
def emul(tmp_path_factory):
    from biggan_tensorflow_b200._lib import parse_header
    so = str(tmp_path_factory.mktemp("host_emul") / "rgb_out_host.so")
    src = os.path.join(ROOT, "biggan_tensorflow_b200", "csrc", "rgb_out.cu")
    subprocess.run(["g++", "-std=c++20", "-O1", "-pthread", "-DBG_HOST_EMULATION", "-x", "c++", "-shared", "-fPIC", src, "-o", so], check=True)
    lib = ctypes.CDLL(so)
    for name, (res, args) in parse_header().items():
        if hasattr(lib, name):
            fn = getattr(lib, name)
            fn.restype, fn.argtypes = res, args
    return lib


def bf16_buf(a):
    """float array -> (16-byte aligned int16 buffer of bf16 bit patterns, the rounded values as float64)"""
    t = torch.from_numpy(np.ascontiguousarray(a, dtype=np.float32)).bfloat16()
    return t.view(torch.int16).numpy().copy(), t.double().numpy()


def from_bf16(buf):
    return torch.from_numpy(buf).view(torch.bfloat16).double().numpy()


def ptr(a):
    return None if a is None else a.ctypes.data_as(ctypes.c_void_p)


def tapsum_ref(Y, reflect):
    """Y [B,H,W,32] torch float64 -> [B,H,W,3]: sum_{a,c} Y[b, R(i+a-1), R(j+c-1), (a*3+c)*3+n]"""
    Yp = torch.nn.functional.pad(Y.permute(0, 3, 1, 2), (1, 1, 1, 1), mode="reflect" if reflect else "constant").permute(0, 2, 3, 1)
    B, H, W = Y.shape[:3]
    out = 0
    for a in range(3):
        for c in range(3):
            out = out + Yp[:, a:a + H, c:c + W, (a * 3 + c) * 3:(a * 3 + c) * 3 + 3]
    return out


CASES = [(2, 8, 8, 16), (1, 5, 30, 16), (1, 3, 31, 8), (2, 2, 2, 16), (1, 4, 61, 16), (1, 2, 128, 16), (1, 9, 3, 24)]


@pytest.mark.parametrize("act", [ACT_TANH, ACT_NONE, ACT_RELU])
@pytest.mark.parametrize("reflect", [1, 0])
@pytest.mark.parametrize("B,H,W,Co", CASES)
def test_tapsum_forward_and_adjoint(emul, B, H, W, Co, reflect, act):
    rng = np.random.default_rng(B * 1000 + H * 10 + W)
    ybuf, yv = bf16_buf(rng.standard_normal((B, H, W, 32)) * 0.5)
    bias = rng.standard_normal(3).astype(np.float32)
    out = np.full((B, H, W, Co), 0x7fc0, dtype=np.int16)           # NaN patterns: every element must be written
    assert emul.bg_rgb_tapsum_fwd(ptr(ybuf), ptr(bias), ptr(out), B, H, W, Co, PAD_REFLECT if reflect else PAD_ZERO, act, None) == 0
    pre = tapsum_ref(torch.from_numpy(yv), reflect) + torch.from_numpy(bias.astype(np.float64))
    want = torch.tanh(pre) if act == ACT_TANH else (torch.relu(pre) if act == ACT_RELU else pre)
    got = from_bf16(out)
    assert np.all(got[..., 3:] == 0.0)
    assert float(np.abs(got[..., :3] - want.numpy()).max()) <= 6e-3 * max(1.0, float(want.abs().max()))

    # adjoint: d pre = d_out * act'(y) from the SAME bf16 y the kernel reads, then autograd through the linear tap sum
    dbuf, dv = bf16_buf(rng.standard_normal((B, H, W, Co)))
    yq = from_bf16(out)
    if act == ACT_TANH:
        dpre = dv[..., :3] * (1.0 - yq[..., :3] ** 2)
    elif act == ACT_RELU:
        dpre = dv[..., :3] * (yq[..., :3] > 0)
    else:
        dpre = dv[..., :3]
    Yt = torch.zeros(B, H, W, 32, dtype=torch.float64, requires_grad=True)
    (tapsum_ref(Yt, reflect) * torch.from_numpy(dpre)).sum().backward()
    dY = np.full((B, H, W, 32), 0x7fc0, dtype=np.int16)
    assert emul.bg_rgb_tapsum_bwd(ptr(out), ptr(dbuf), ptr(dY), B, H, W, Co, PAD_REFLECT if reflect else PAD_ZERO, act, None) == 0
    gotd = from_bf16(dY)
    wantd = Yt.grad.numpy()
    assert np.all(gotd[..., 27:] == 0.0) and np.all(wantd[..., 27:] == 0.0)
    assert float(np.abs(gotd - wantd).max()) <= 6e-3 * max(1.0, float(np.abs(wantd).max()))


@pytest.mark.parametrize("Ci,c_pad", [(16, 16), (96, 16), (8, 8)])
def test_stack_and_unstack_weights(emul, Ci, c_pad):
    rng = np.random.default_rng(Ci)
    wbuf, wv = bf16_buf(rng.standard_normal((3, 3, Ci, c_pad)))
    wall = np.full((Ci, 32), 0x7fc0, dtype=np.int16)
    wall_t = np.full((32, Ci), 0x7fc0, dtype=np.int16)
    assert emul.bg_rgb_stack_weights(ptr(wbuf), ptr(wall), ptr(wall_t), Ci, c_pad, None) == 0
    a, t = from_bf16(wall), from_bf16(wall_t)
    want = np.zeros((Ci, 32))
    for tap in range(9):
        for n in range(3):
            want[:, tap * 3 + n] = wv[tap // 3, tap % 3, :, n]
    assert np.array_equal(a, want) and np.array_equal(t, want.T)
    dwall = rng.standard_normal((Ci, 32)).astype(np.float32)
    dwn = np.full((3, 3, Ci, c_pad), np.nan, dtype=np.float32)
    assert emul.bg_rgb_unstack_wgrad(ptr(dwall), ptr(dwn), Ci, c_pad, None) == 0
    for tap in range(9):
        for n in range(c_pad):
            exp = dwall[:, tap * 3 + n] if n < 3 else np.zeros(Ci, np.float32)
            assert np.array_equal(dwn[tap // 3, tap % 3, :, n], exp)


def test_shape_errors(emul):
    z = np.zeros(64, np.int16)
    assert emul.bg_rgb_tapsum_fwd(ptr(z), None, ptr(z), 1, 1, 8, 16, PAD_REFLECT, ACT_NONE, None) != 0      # H < 2
    assert emul.bg_rgb_tapsum_fwd(ptr(z), None, ptr(z), 1, 4, 4, 12, PAD_REFLECT, ACT_NONE, None) != 0     # Co % 8
