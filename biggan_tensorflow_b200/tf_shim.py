"""`tf.variable_scope` / `tf.get_variable` shim: name -> device buffer.

The reference creates parameters by side effect through TF variable scopes (ops.py:50, :88, :127,
:156, :620-621, :722); the names are the key for loading identical weights on both sides
(SURVEY.md appendix A.3).  Initialisers are deterministic, keyed by the variable's full name.
"""
import contextlib
import zlib

import numpy as np
import torch

from . import _lib
from .tensor import Tensor, device, stream


def name_seed(name):
    return zlib.crc32(name.encode("utf-8")) & 0xFFFFFFFF


class Initializer:
    def __init__(self, kind, mean=0.0, stddev=1.0, value=0.0):
        self.kind, self.mean, self.stddev, self.value = kind, mean, stddev, value

    def __call__(self, name, shape):
        shape = tuple(int(s) for s in shape)
        n = int(np.prod(shape)) if len(shape) else 1
        if self.kind == "const":
            return np.full(shape, self.value, dtype=np.float32)
        rng = np.random.Generator(np.random.PCG64(name_seed(name)))
        if self.kind == "normal":
            return (rng.standard_normal(n) * self.stddev + self.mean).astype(np.float32).reshape(shape)
        if self.kind == "trunc_normal":          # values beyond 2 sigma are re-drawn (TF semantics)
            x = rng.standard_normal(n)
            bad = np.abs(x) > 2.0
            while bad.any():
                x[bad] = rng.standard_normal(int(bad.sum()))
                bad = np.abs(x) > 2.0
            return (x * self.stddev + self.mean).astype(np.float32).reshape(shape)
        raise ValueError(self.kind)


def truncated_normal_initializer(mean=0.0, stddev=1.0):
    return Initializer("trunc_normal", mean=mean, stddev=stddev)


def random_normal_initializer(mean=0.0, stddev=1.0):
    return Initializer("normal", mean=mean, stddev=stddev)


def constant_initializer(value=0.0):
    return Initializer("const", value=value)


class Variable(Tensor):
    __slots__ = ("trainable",)

    def __init__(self, data, name, trainable):
        super().__init__(data, requires_grad=bool(trainable), name=name)
        self.trainable = trainable


class VariableStore:
    def __init__(self):
        self.vars = {}
        self.scope = []
        self.sn_cache = {}       # per update-phase cache: one spectral_norm evaluation per variable
        self.sn_grouped = {}     # u name -> cache key of the weights a grouped spectral norm already advanced this phase
        self.cbn_cache = {}      # batch-norm scope name -> (gamma, beta) computed up front (ops.precompute_cbn_projections)
        self.arenas = {}
        self.regularizers = {}   # full variable name -> regularizer (tf REGULARIZATION_LOSSES: one entry per variable)
        self.default_names = {}  # parent scope -> {base: count} for variable_scope(None, default_name=...)

    def full_name(self, name=None):
        parts = list(self.scope)
        if name is not None:
            parts.append(name)
        return "/".join(parts)

    def get_variable(self, name, shape, initializer=None, trainable=True, dtype=None, regularizer=None):
        full = self.full_name(name)
        if regularizer is not None:
            self.regularizers[full] = regularizer
        v = self.vars.get(full)
        shape = tuple(int(s) for s in shape)
        if v is None:
            init = initializer or constant_initializer(0.0)
            host = torch.from_numpy(init(full, shape))
            v = Variable(host.to(device()), full, trainable)
            self.vars[full] = v
        if tuple(v.data.shape) != shape:
            raise ValueError("variable %s has shape %s, requested %s" % (full, tuple(v.data.shape), shape))
        return v

    def begin_phase(self):
        self.sn_cache = {}
        self.sn_grouped = {}
        self.cbn_cache = {}

    def trainable_variables(self, substr):
        return [v for k, v in sorted(self.vars.items()) if v.trainable and substr in k]

    def load(self, params):
        """params: name -> numpy array (e.g. exported from the oracle store)."""
        for k, a in params.items():
            a = np.asarray(a, dtype=np.float32)
            if k in self.vars:
                self.vars[k].data.copy_(torch.from_numpy(a).reshape(self.vars[k].data.shape))
            else:
                trainable = not (k.endswith("/u") or "pop_" in k or "moving_" in k)
                self.vars[k] = Variable(torch.from_numpy(a.copy()).to(device()), k, trainable)

    def export(self):
        torch.cuda.synchronize()
        return {k: v.data.detach().float().cpu().numpy().copy() for k, v in self.vars.items()}

    def pack(self, substr, multiple=1):
        """Move the trainable variables containing `substr` into one flat fp32 arena (params, grads)
        so that Adam/EMA and the gradient all-reduce are single launches over contiguous memory.
        `multiple`: round the arena length up to a multiple of it (equal 16-byte aligned shards per rank)."""
        vs = self.trainable_variables(substr)
        offs, total = [], 0
        for v in vs:
            offs.append(total)
            total += (v.data.numel() + 3) // 4 * 4           # keep every slice 16-byte aligned
        total = (total + multiple - 1) // multiple * multiple
        p = torch.zeros(total, dtype=torch.float32, device=device())
        g = torch.zeros(total, dtype=torch.float32, device=device())
        for v, o in zip(vs, offs):
            n = v.data.numel()
            p[o:o + n].copy_(v.data.reshape(-1))
            v.data = p[o:o + n].view(v.data.shape)
            v.grad_view = g[o:o + n].view(v.data.shape)
        self.arenas[substr] = {"params": p, "grads": g, "vars": vs, "offsets": offs}
        return self.arenas[substr]


_STORE = VariableStore()


def get_store():
    return _STORE


def set_store(store):
    global _STORE
    _STORE = store
    return store


@contextlib.contextmanager
def variable_scope(name=None, reuse=None, default_name=None, name_or_scope=None, **_unused):
    """tf.variable_scope: pushes `name`; `reuse` is accepted for signature parity (get_variable
    always returns the existing variable of that name, like reuse=tf.AUTO_REUSE).  With name None the
    `default_name` is made unique within the parent scope the way TF does ('prelu', 'prelu_1', ...; ops.py:533);
    the counters of everything below a scope are reset when the scope is left, so a second instantiation of a
    network names its default scopes identically."""
    st = _STORE
    if name is None:
        name = name_or_scope
    if name is None:
        d = st.default_names.setdefault(st.full_name(), {})
        k = d.get(default_name, 0)
        d[default_name] = k + 1
        name = default_name if k == 0 else "%s_%d" % (default_name, k)
    st.scope.append(name)
    try:
        yield st
    finally:
        full = st.full_name()
        for key in [key for key in st.default_names if key == full or key.startswith(full + "/")]:
            del st.default_names[key]
        st.scope.pop()


def get_variable(name, shape=None, initializer=None, trainable=True, dtype=None, regularizer=None):
    return _STORE.get_variable(name, shape, initializer, trainable, dtype, regularizer)


def current_scope_name():
    return _STORE.full_name()
