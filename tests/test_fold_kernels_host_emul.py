"""CPU execution check of the per-tap dy-fold kernels (`csrc/fold_kernels.cuh`: the small-map dgrad path of reflect-padded 3x3 convs,
ops.py:65-84): the SAME kernel source compiled for the host with `-DBG_HOST_EMULATION` (tests/host_emul/cuda_host_emul.h), launched
through test-only wrappers (tests/host_emul/fold_host.cu), against the literal enumeration of tests/test_fold_identities.py.
The kernels on hardware: tests/test_conv_epilogue_gpu.py."""
import ctypes
import os
import subprocess

import numpy as np
import pytest
import torch

from test_fold_identities import fold_dy

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def emul(tmp_path_factory):
    so = str(tmp_path_factory.mktemp("host_emul") / "fold_host.so")
    src = os.path.join(ROOT, "tests", "host_emul", "fold_host.cu")
    subprocess.run(["g++", "-std=c++20", "-O1", "-pthread", "-DBG_HOST_EMULATION", "-x", "c++", "-shared", "-fPIC", src, "-o", so], check=True)
    lib = ctypes.CDLL(so)
    lib.test_reflect_dyfold.restype = ctypes.c_int
    lib.test_reflect_dyfold.argtypes = [ctypes.c_void_p, ctypes.c_void_p] + [ctypes.c_int] * 5
    return lib


@pytest.mark.parametrize("B,H,W,C,use4", [(3, 4, 4, 16, 1), (20, 4, 4, 40, 1), (3, 4, 4, 16, 0), (2, 3, 3, 8, 0), (5, 2, 2, 24, 0), (2, 4, 3, 16, 0),
                                          (1, 2, 4, 8, 0)])
def test_dy_fold_kernels(emul, B, H, W, C, use4):
    rng = np.random.default_rng(B * 100 + H * 10 + W)
    dy = torch.from_numpy(rng.standard_normal((B, H, W, C)).astype(np.float32)).bfloat16()
    src = dy.view(torch.int16).numpy().copy()
    dst = np.full((B, 9 * H, W, C), 0x7fc0, dtype=np.int16)        # NaN patterns: every element must be written
    assert emul.test_reflect_dyfold(src.ctypes.data, dst.ctypes.data, B, H, W, C, use4) == 0
    got = torch.from_numpy(dst).view(torch.bfloat16).double()                                   # [B, 9H, W, C]
    want = fold_dy(dy.double()).permute(1, 0, 2, 3, 4).reshape(B, 9 * H, W, C)                  # tap t at rows t*H .. t*H+H-1
    assert bool(torch.isfinite(got).all())
    # sums of at most four bf16 values, accumulated in fp32 and rounded to bf16 once
    assert float((got - want).abs().max()) <= 2.0 ** -8 * float(want.abs().max())
    assert torch.equal(got, want.bfloat16().double())
