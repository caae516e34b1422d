"""Device tensor handle + reverse-mode tape.

torch is used ONLY as the device-memory allocator / stream provider (tensor plumbing); every
arithmetic operation on the hot path is a kernel of libbiggan_b200.so.  The tape replaces
`tf.gradients` (BigGAN.py:923-930 via utils.py:258-260): ops append a backward closure, and
`backward()` replays them in reverse.
"""
import contextlib

import torch

from . import _lib

F32, BF16 = torch.float32, torch.bfloat16


def bg_dtype(t):
    dt = t.dtype if hasattr(t, "dtype") else t
    if dt == torch.float32:
        return _lib.BG_F32
    if dt == torch.bfloat16:
        return _lib.BG_BF16
    raise ValueError("unsupported dtype %s" % dt)


def stream():
    return torch.cuda.current_stream().cuda_stream


def device():
    return torch.device("cuda", torch.cuda.current_device())


def empty(shape, dtype=F32):
    return torch.empty(tuple(int(s) for s in shape), dtype=dtype, device=device())


_ws_cache = {}


def workspace(nbytes, slot=0):
    """Grow-only scratch buffers (stream-ordered reuse; one compute stream)."""
    nbytes = int(nbytes)
    key = (torch.cuda.current_device(), slot)
    buf = _ws_cache.get(key)
    if buf is None or buf.numel() < nbytes:
        buf = torch.empty(max(nbytes, 1 << 20), dtype=torch.uint8, device=device())
        _ws_cache[key] = buf
    return buf


class Tensor:
    """NHWC activation / parameter handle.  `.data` is the device buffer (a torch tensor)."""
    __slots__ = ("data", "grad", "requires_grad", "name", "grad_view", "grad_set", "grad_owned", "aux", "relu_out", "n_contribs",
                 "n_masked")

    def __init__(self, data, requires_grad=False, name=None):
        self.data = data
        self.grad = None
        self.requires_grad = requires_grad
        self.name = name
        self.grad_view = None      # preallocated slice of a flat gradient arena (parameters)
        self.grad_set = False
        self.grad_owned = True     # False while .grad aliases a buffer another tensor may also hold
        self.aux = None
        # ReLU-backward fusion (ops._conv_op): `relu_out` marks the output of a conv with a fused ReLU; every gradient
        # contribution is counted, and those a consumer's dgrad already multiplied by the mask (data > 0) separately -- the
        # producer skips its own mask pass exactly when all contributions arrived masked
        self.relu_out = False
        self.n_contribs = 0
        self.n_masked = 0

    @property
    def shape(self):
        return tuple(self.data.shape)

    @property
    def dtype(self):
        return self.data.dtype

    def get_shape(self):           # TF-style accessors used by the reference call sites
        return list(self.data.shape)

    def numel(self):
        return self.data.numel()

    def numpy(self):
        return self.data.detach().float().cpu().numpy()


class _Tape:
    def __init__(self):
        self.nodes = []
        self.enabled = True


TAPE = _Tape()


@contextlib.contextmanager
def no_grad():
    prev = TAPE.enabled
    TAPE.enabled = False
    try:
        yield
    finally:
        TAPE.enabled = prev


def recording(*inputs):
    return TAPE.enabled and any(t is not None and t.requires_grad for t in inputs)


def record(fn):
    TAPE.nodes.append(fn)


def clear_tape():
    TAPE.nodes = []


def _add_into_grad(t, g):
    """t.grad += g without ever mutating a borrowed (shared) buffer."""
    if t.grad_owned:
        _lib.call("bg_add", t.grad, g, t.grad, t.grad.numel(), bg_dtype(t.grad), stream())
    else:
        out = torch.empty_like(t.grad)
        _lib.call("bg_add", t.grad, g, out, out.numel(), bg_dtype(out), stream())
        t.grad = out
        t.grad_owned = True


def grad_target(t, dtype=None):
    """Buffer a backward kernel should write d(loss)/d(t) into, and a finisher to call afterwards.
    `dtype` overrides the gradient storage type (weight gradients are always fp32).

    First contribution: written in place (into the parameter's arena slice when it has one).
    Later contributions: written to a temporary and accumulated with bg_add."""
    t.n_contribs += 1
    if t.grad_view is not None and not t.grad_set:
        t.grad_set = True
        t.grad = t.grad_view
        t.grad_owned = True
        return t.grad_view, None
    if t.grad is None:
        t.grad = torch.empty_like(t.data, dtype=dtype or t.data.dtype)
        t.grad_set = True
        t.grad_owned = True
        return t.grad, None
    tmp = torch.empty_like(t.grad)
    return tmp, (lambda: _add_into_grad(t, tmp))


def accumulate_grad(t, g, own=False):
    """t.grad (+)= g.  With own=True the buffer g may be adopted as t.grad and mutated later;
    otherwise it is only borrowed (copied on the next accumulation)."""
    t.n_contribs += 1
    if t.grad_view is not None and not t.grad_set:
        t.grad_view.copy_(g.reshape(t.grad_view.shape))
        t.grad_set = True
        t.grad = t.grad_view
        t.grad_owned = True
    elif t.grad is None:
        t.grad = g.reshape(t.data.shape)
        t.grad_set = True
        t.grad_owned = own
    else:
        _add_into_grad(t, g.reshape(t.grad.shape))


def backward(loss_grads):
    """loss_grads: list of (tensor, grad buffer) seeds.  Replays the tape in reverse."""
    for t, g in loss_grads:
        accumulate_grad(t, g, own=True)
    nodes = TAPE.nodes
    TAPE.nodes = []
    for fn in reversed(nodes):
        fn()
