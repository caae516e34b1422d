"""Shared helpers of the GPU parity tests: run the product op (CUDA, through the C ABI) and the
oracle op (torch CPU float64) on the same seeded inputs and identical name-seeded variables."""
import numpy as np
import torch

from oracle import ref_ops as R


ABS_FLOOR = 1e-6   # tensors whose reference is identically ~0 (e.g. d/d f_conv bias: softmax is shift invariant)


def rel_err(a, b):
    """max |a-b| / max|ref|  (the fp32-mode metric)."""
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    denom = max(float(np.abs(b).max()), ABS_FLOOR)
    return float(np.abs(a - b).max()) / denom


def rel_l2(a, b):
    """||a-b||_2 / ||ref||_2  (the bf16-mode metric: single elements may flip a ReLU mask under bf16
    rounding, which moves one gradient term by O(1) without being a kernel error)."""
    a = np.asarray(a, dtype=np.float64).ravel()
    b = np.asarray(b, dtype=np.float64).ravel()
    denom = max(float(np.linalg.norm(b)), ABS_FLOOR * np.sqrt(b.size))
    return float(np.linalg.norm(a - b)) / denom


def assert_close(a, b, tol, what=""):
    """tol < 1e-3: fp32 mode, max-abs error relative to max|ref| <= tol (north_star: 1e-4).
    tol >= 1e-3: bf16 mode, relative L2 error <= tol (north_star: 2e-2) and max-abs error <= 10*tol*max|ref|."""
    assert np.isfinite(np.asarray(a, dtype=np.float64)).all(), what + ": non-finite values"
    if tol < 1e-3:
        e = rel_err(a, b)
        assert e <= tol, "%s: max-abs error relative to max|ref| = %.3e > %.1e" % (what, e, tol)
    else:
        e = rel_l2(a, b)
        assert e <= tol, "%s: relative L2 error = %.3e > %.1e" % (what, e, tol)
        m = rel_err(a, b)
        assert m <= max(10 * tol, 0.5), "%s: max-abs error relative to max|ref| = %.3e" % (what, m)


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
