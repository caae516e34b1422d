"""Shared helpers of the GPU parity tests: run the product op (CUDA, through the C ABI) and the
oracle op (torch CPU float64) on the same seeded inputs and identical name-seeded variables."""
import numpy as np
import torch

from oracle import ref_ops as R


def rel_err(a, b):
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    denom = max(float(np.abs(b).max()), 1e-30)
    return float(np.abs(a - b).max()) / denom


def assert_close(a, b, tol, what=""):
    e = rel_err(a, b)
    assert e <= tol, "%s: max-abs error relative to max|ref| = %.3e > %.1e" % (what, e, tol)


def new_product_store(dtype):
    from biggan_tensorflow_b200 import ops
    from biggan_tensorflow_b200.tf_shim import VariableStore, set_store
    from biggan_tensorflow_b200.tensor import clear_tape
    st = set_store(VariableStore())
    ops.set_compute_dtype(dtype)
    ops.dist.world_size = 1
    ops.dist.all_reduce_sum = lambda b: b
    clear_tape()
    return st


def dev_tensor(a, dtype=torch.float32, requires_grad=True):
    from biggan_tensorflow_b200.tensor import Tensor
    return Tensor(torch.from_numpy(np.ascontiguousarray(a)).cuda().to(dtype), requires_grad=requires_grad)


def seed_backward(y, dy_np):
    from biggan_tensorflow_b200.tensor import backward
    g = torch.from_numpy(np.ascontiguousarray(dy_np)).cuda().to(y.dtype)
    backward([(y, g)])
    torch.cuda.synchronize()


def ref_grads(y_ref, dy_np, wrt):
    dy = torch.as_tensor(dy_np, dtype=y_ref.dtype)
    return torch.autograd.grad((y_ref * dy).sum(), wrt, allow_unused=True)


TOL = {torch.float32: 1e-4, torch.bfloat16: 2e-2}   # north_star tolerances (relative)
