"""DiffAugment entry point of the reference (DiffAugment_tf.py:8-17) on libbiggan_b200.so.

`DiffAugment(x, policy='', channels_first=False)` keeps the reference signature.  The per-sample
random magnitudes / integer offsets are drawn HOST-side in exactly the ranges of the reference
(DiffAugment_tf.py:21,:27,:34,:44-45,:57-58) from `random_source` (a numpy Generator), or taken from
a queue pushed with `push_randoms` (parity tests feed both implementations the same draws -- TF's
Philox streams are not reproducible, so randomness is an input).
"""
import numpy as np
import torch

from . import _lib
from ._lib import call
from .tensor import F32, Tensor, accumulate_grad, bg_dtype, device, empty, record, recording, stream

random_source = np.random.Generator(np.random.PCG64(4))
_queue = []

_FLAG = {"color": 1, "translation": 2, "cutout": 4}
AUGMENT_FNS = {                                            # DiffAugment_tf.py:69-73
    'color': ['rand_brightness', 'rand_saturation', 'rand_contrast'],
    'translation': ['rand_translation'],
    'cutout': ['rand_cutout'],
}


def translation_shift(size, ratio=0.125):
    return int(size * ratio + 0.5)                         # DiffAugment_tf.py:43


def cutout_size(size, ratio=0.5):
    return int(size * ratio + 0.5)                         # DiffAugment_tf.py:56


def draw_randoms(rng, B, H, W):
    sh, sw = translation_shift(H), translation_shift(W)
    cs0, cs1 = cutout_size(H), cutout_size(W)
    return {
        "brightness": rng.random(B, dtype=np.float32),
        "saturation": rng.random(B, dtype=np.float32),
        "contrast": rng.random(B, dtype=np.float32),
        "translation_x": rng.integers(-sh, sh + 1, B).astype(np.int32),          # [-shift, shift]
        "translation_y": rng.integers(-sw, sw + 1, B).astype(np.int32),
        "cutout_x": rng.integers(0, H + (1 - cs0 % 2), B).astype(np.int32),      # [0, H-1+(1-cs%2)]
        "cutout_y": rng.integers(0, W + (1 - cs1 % 2), B).astype(np.int32),
    }


def push_randoms(rnd):
    """Queue explicit draws for the next DiffAugment call (FIFO)."""
    _queue.append(rnd)


def _upload(rnd, B):
    f = np.stack([rnd["brightness"], rnd["saturation"], rnd["contrast"]]).astype(np.float32)
    i = np.stack([rnd["translation_x"], rnd["translation_y"], rnd["cutout_x"], rnd["cutout_y"]]).astype(np.int32)
    assert f.shape == (3, B) and i.shape == (4, B)
    return (torch.from_numpy(f).to(device(), non_blocking=True),
            torch.from_numpy(i).to(device(), non_blocking=True))


def _apply(x, flags, rnd_f, rnd_i):
    B, H, W, C = x.shape
    y = Tensor(empty(x.shape, x.dtype), requires_grad=x.requires_grad)
    mean_ws = empty([B], F32)
    call("bg_diffaug_fwd", x.data, y.data, rnd_f, rnd_i, flags, B, H, W, C, mean_ws, bg_dtype(x), stream())
    if recording(x):
        def bw():
            if y.grad is None:
                return
            dx = torch.empty_like(x.data)
            mws = empty([B], F32)
            call("bg_diffaug_bwd", y.grad, dx, rnd_f, rnd_i, flags, B, H, W, C, mws, bg_dtype(x), stream())
            accumulate_grad(x, dx, own=True)
        record(bw)
    return y


def DiffAugment(x, policy='', channels_first=False):
    """DiffAugment_tf.py:8-17.  x: NHWC Tensor."""
    if policy:
        if channels_first:
            raise ValueError("channels_first=True is not supported (the hot path is NHWC)")
        parts = policy.split(',')
        for p in parts:
            if p not in AUGMENT_FNS:
                raise KeyError(p)
        from . import ops as _ops
        c_pad, c_log = x.shape[-1], _ops.c_real(x)
        if c_pad != c_log:                       # zero-padded image tensor: augment the logical channels only
            x = _ops.slice_channels(x, c_log)
        B, H, W, C = x.shape
        rnd = _queue.pop(0) if _queue else draw_randoms(random_source, B, H, W)
        rnd_f, rnd_i = _upload(rnd, B)
        # fuse maximal runs that follow the canonical order color -> translation -> cutout
        i = 0
        while i < len(parts):
            flags, last = 0, 0
            while i < len(parts) and _FLAG[parts[i]] > last:
                flags |= _FLAG[parts[i]]
                last = _FLAG[parts[i]]
                i += 1
            x = _apply(x, flags, rnd_f, rnd_i)
        if c_pad != c_log:
            x = _ops.pad_channels(x, c_pad)
    return x
