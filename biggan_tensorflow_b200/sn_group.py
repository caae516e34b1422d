"""Grouped spectral norm (ops.py:718-747 for every weight of a network at once).

`ops.spectral_norm` evaluates one matrix per call (3+1 launches forward, 2 backward); a BigGAN-128 step needs it
for 64 matrices, four times -- ~640 tiny launches whose cost is pure latency and does not shrink with the per-GPU
batch.  The Trainer therefore registers every spectral_norm call made while the networks are first built, gives
each its persistent output / gradient buffers, and runs the whole network through `bg_sn_group_forward`
(4 launches) at the start of an update phase and `bg_sn_group_backward` (2 launches) after the tape replay.
The per-call results are published through the store's phase cache, so the operator surface is unchanged."""
import ctypes

import torch

from . import _lib
from ._lib import call, query
from .tensor import F32, Tensor, bg_dtype, device, stream


class SnItem(ctypes.Structure):
    """mirror of bg_sn_item"""
    _fields_ = [(n, ctypes.c_void_p) for n in ("W", "u", "v", "sigma", "out", "out_t", "dWn", "dW")] + \
               [(n, ctypes.c_int) for n in ("K", "C", "taps", "rows", "r_pad", "c_pad", "dtype", "g_ld")]


class SnGroup:
    def __init__(self, entries):
        """entries: registry records of ops.spectral_norm (dicts with w, u, taps, rows, r_pad, c_pad, out_dtype, want_t, key)."""
        self.entries = entries
        n = len(entries)
        # one record per weight: two records sharing a `u` (the same weight requested in two layouts) would both advance u and
        # both write dW within one launch
        names = [e["u"].name for e in entries]
        assert len(set(names)) == len(names), "spectral_norm registry holds two layouts of one weight: %s" % sorted(
            n_ for n_ in set(names) if names.count(n_) > 1)
        self.items = (SnItem * n)()
        dev = device()
        for i, e in enumerate(entries):
            w, u = e["w"], e["u"]
            C = w.shape[-1]
            K = w.numel() // C
            shape = list(w.shape[:-2]) + [e["r_pad"], e["c_pad"]]
            e["v"] = torch.empty(K, dtype=F32, device=dev)
            e["sigma"] = torch.empty(1, dtype=F32, device=dev)
            e["out"] = torch.zeros(shape, dtype=e["out_dtype"], device=dev)
            e["out_t"] = torch.zeros(shape, dtype=e["out_dtype"], device=dev) if e["want_t"] else None
            e["dWn"] = torch.zeros(shape, dtype=F32, device=dev)
            assert w.grad_view is not None, "pack() the variables before building the SN group"
            it = self.items[i]
            it.W, it.u, it.v, it.sigma = w.data.data_ptr(), u.data.data_ptr(), e["v"].data_ptr(), e["sigma"].data_ptr()
            it.out = e["out"].data_ptr()
            it.out_t = e["out_t"].data_ptr() if e["out_t"] is not None else None
            it.dWn, it.dW = e["dWn"].data_ptr(), w.grad_view.data_ptr()
            it.K, it.C, it.taps, it.rows, it.r_pad, it.c_pad = K, C, e["taps"], e["rows"], e["r_pad"], e["c_pad"]
            it.dtype, it.g_ld = bg_dtype(e["out_dtype"]), e["c_pad"]
        nbytes = _lib.lib.bg_sn_group_plan_bytes(ctypes.byref(self.items), n)
        self.plan = torch.empty(int(nbytes), dtype=torch.uint8, device=dev)
        self.totals = (ctypes.c_int * 8)()
        rc = _lib.lib.bg_sn_group_plan(ctypes.byref(self.items), n, self.plan.data_ptr(), int(nbytes), self.totals)
        if rc != 0:
            raise _lib.BgError("bg_sn_group_plan failed: " + _lib.lib.bg_last_error().decode())

    def forward(self, store):
        """Power iteration + scale for every matrix; publishes the normalised weights in the phase cache."""
        call("bg_sn_group_forward", self.plan, ctypes.addressof(self.totals), stream(), tag="sn_group_forward")
        self.live = []
        for e in self.entries:
            w = e["w"]
            store.sn_grouped[e["u"].name] = e["key"]     # a later call asking for another layout of this weight must not re-run the power iteration
            wn = Tensor(e["out"], requires_grad=w.requires_grad)
            wn.aux = {"w_t": e["out_t"], "sigma": e["sigma"], "v_hat": e["v"], "u_hat": e["u"].data}
            if w.requires_grad:
                wn.grad_view = e["dWn"]          # conv wgrad writes its fp32 result straight into the persistent buffer
            store.sn_cache[e["key"]] = wn
            self.live.append((e, wn))

    def backward(self):
        """dW for every trained matrix (call after the tape replay)."""
        for e, wn in self.live:
            if wn.requires_grad and not wn.grad_set:
                e["dWn"].zero_()                 # weight that received no gradient this phase
        call("bg_sn_group_backward", self.plan, ctypes.addressof(self.totals), stream(), tag="sn_group_backward")
        for e, wn in self.live:
            w = e["w"]
            if w.requires_grad:
                w.grad, w.grad_set, w.grad_owned = w.grad_view, True, True
