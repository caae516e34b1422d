"""TEST INFRASTRUCTURE ONLY -- an eager stand-in for the slice of the TensorFlow-1.x API that the
reference (david-jk/BigGAN-Tensorflow: ops.py, DiffAugment_tf.py, utils.py regularisers / class loss,
BigGAN.py generator() / discriminator()) calls on the hot path.

Why it exists: TensorFlow 1.x cannot be installed in this image, so the reference cannot be executed as
shipped.  With this module registered as ``tensorflow`` (``install()``), the reference's OWN Python source
is imported unmodified from /root/reference and executed; `oracle/make_reference_golden.py` does that and
writes `tests/golden/ref_*.npz`.  Those vectors pin everything the reference source decides -- padding
arithmetic, power-iteration order, stop_gradients, variable names and shapes, block wiring, z splitting,
attention wiring, loss formulas, DiffAugment index math -- to the reference itself.  What they do NOT pin
is TensorFlow's kernels: each primitive below is written from TF-1.x's documented semantics (the doc string
of every function states them) on torch-CPU float64, deliberately NOT sharing code with oracle/ref_ops.py
(e.g. conv2d_transpose is literally "the gradient of conv2d with respect to its input", REFLECT padding is
an index gather, scatter_nd accumulates).

Session model (documented choice): the reference builds one graph and runs an update phase with one
``sess.run``.  Inside a run, variable READS see the value from before the run and ``assign`` ops take
effect when the run ends (reads are not control-dependent on the assigns; two instantiations of the
discriminator in one run therefore see the same spectral-norm ``u``).  Here: ``assign`` is deferred until
``end_run()``.

Nothing under biggan_tensorflow_b200/ imports this file.
"""
import contextlib
import sys
import types

import numpy as np
import torch
import torch.nn.functional as F

FLOAT = torch.float64          # what tf.float32 maps to while generating golden vectors


# ---------------------------------------------------------------------------------------------
# dtypes / shapes / tensors
# ---------------------------------------------------------------------------------------------

class DType:
    def __init__(self, name):
        self.name = name

    def torch(self):
        return {"float32": FLOAT, "int32": torch.int64, "bool": torch.bool}[self.name]

    def __repr__(self):
        return "tf." + self.name


float32 = DType("float32")
int32 = DType("int32")


class TensorShape(tuple):
    def as_list(self):
        return [int(d) for d in self]

    def __getitem__(self, i):
        r = tuple.__getitem__(self, i)
        return TensorShape(r) if isinstance(i, slice) else r


def _raw(x):
    if isinstance(x, Tensor):
        return x.t
    return x


def _t(x, like=None):
    """To a torch tensor (python scalars / lists / numpy / Tensor)."""
    if isinstance(x, Tensor):
        return x.t
    if isinstance(x, torch.Tensor):
        return x
    if isinstance(x, (list, tuple)) and any(isinstance(e, Tensor) for e in x):
        return torch.stack([_t(e) for e in x])
    if isinstance(x, (bool, int)) and like is not None and not like.is_floating_point():
        return torch.as_tensor(x, dtype=like.dtype)
    if isinstance(x, (int, float)):
        return torch.as_tensor(x, dtype=FLOAT if like is None or like.is_floating_point() or isinstance(x, float) else like.dtype)
    a = np.asarray(x)
    return torch.as_tensor(a, dtype=FLOAT if a.dtype.kind == "f" else torch.int64)


class Tensor:
    """Eager value.  `.t` is the torch tensor (autograd-tracked when it depends on a trainable variable)."""

    def __init__(self, t, name=None):
        self._t = t
        self.name = name

    @property
    def t(self):
        return self._t

    @property
    def shape(self):
        return TensorShape(self.t.shape)

    def get_shape(self):
        return self.shape

    @property
    def dtype(self):
        return float32 if self.t.is_floating_point() else int32

    def _bin(self, other, fn, rev=False):
        a = self.t
        b = _t(other, like=a)
        if a.is_floating_point() != b.is_floating_point():
            # TF would refuse int*float; the reference only mixes python scalars, converted above
            raise TypeError("dtype mismatch in eager shim: %s vs %s" % (a.dtype, b.dtype))
        return Tensor(fn(b, a) if rev else fn(a, b))

    def __add__(self, o): return self._bin(o, torch.add)
    def __radd__(self, o): return self._bin(o, torch.add, True)
    def __sub__(self, o): return self._bin(o, torch.sub)
    def __rsub__(self, o): return self._bin(o, torch.sub, True)
    def __mul__(self, o): return self._bin(o, torch.mul)
    def __rmul__(self, o): return self._bin(o, torch.mul, True)
    def __truediv__(self, o): return self._bin(o, torch.true_divide)
    def __rtruediv__(self, o): return self._bin(o, torch.true_divide, True)
    def __floordiv__(self, o): return self._bin(o, lambda a, b: torch.div(a, b, rounding_mode="floor"))
    def __mod__(self, o): return self._bin(o, torch.remainder)      # python/TF floormod semantics
    def __neg__(self): return Tensor(-self.t)
    def __abs__(self): return Tensor(torch.abs(self.t))

    def __getitem__(self, idx):
        if not isinstance(idx, tuple):
            idx = (idx,)
        idx = tuple(int(_raw(i)) if isinstance(i, Tensor) else i for i in idx)
        return Tensor(self.t[idx])

    def __int__(self): return int(self.t)
    def __index__(self): return int(self.t)
    def __float__(self): return float(self.t)

    def __bool__(self):
        raise TypeError("a tf.Tensor has no truth value in graph mode")

    def __iter__(self):
        raise TypeError("tf.Tensor is not iterable in graph mode")

    def numpy(self):
        return self.t.detach().cpu().numpy()


# ---------------------------------------------------------------------------------------------
# graph state: variable scopes, variables, collections, deferred assigns
# ---------------------------------------------------------------------------------------------

class _State:
    def __init__(self):
        self.reset()

    def reset(self):
        self.scope = []                # list of (name, reuse)
        self.vars = {}                 # full name -> Variable (creation order kept)
        self.pending = {}              # full name -> torch value committed by end_run()
        self.pending_sub = {}          # full name -> accumulated assign_sub deltas committed by end_run()
        self.regularizers = []         # (callable, Variable)
        self.default_names = {}        # parent scope -> {base: count}
        self.random_queue = []         # values handed out by tf.random.uniform, in call order
        self.random_log = []           # (shape, minval, maxval, dtype-name, value) of every draw
        self.random_rng = np.random.Generator(np.random.PCG64(1234))
        self.initial_values = None     # optional {full name: ndarray}
        self.init_fn = None            # (full name, shape, kind, value, stddev) -> ndarray


STATE = _State()


def reset_default_graph():
    STATE.reset()


def end_run():
    """End of one sess.run: commit the assigns made during the run."""
    for name, val in STATE.pending.items():
        STATE.vars[name]._set(val)
    for name, delta in STATE.pending_sub.items():
        STATE.vars[name]._set(STATE.vars[name].t.detach() - delta)
    STATE.pending = {}
    STATE.pending_sub = {}


class Variable(Tensor):
    def __init__(self, full, value, trainable):
        self.full = full
        self.name = full + ":0"
        self.trainable = trainable
        self._set(value)

    def _set(self, value):
        v = value.detach().clone()
        if self.trainable:
            v.requires_grad_(True)
        self._t = v

    def assign(self, value):
        v = _t(value).detach()
        assert tuple(v.shape) == tuple(self._t.shape), (self.full, v.shape, self._t.shape)
        STATE.pending[self.full] = v
        return Tensor(v)

    def assign_sub(self, delta):
        """In-place subtraction (what moving_averages.assign_moving_average emits): several of them in one run compose
        additively, each delta having been computed from the pre-run value."""
        d = _t(delta).detach()
        STATE.pending_sub[self.full] = STATE.pending_sub.get(self.full, 0) + d
        return Tensor(d)

    def read_value(self):
        return Tensor(self._t)

    @property
    def op(self):
        return self


class _Scope:
    def __init__(self, name, reuse):
        self._name, self._reuse = name, reuse
        self.name = None

    def __enter__(self):
        parent_reuse = STATE.scope[-1][1] if STATE.scope else False
        reuse = self._reuse if self._reuse is not None else parent_reuse
        if self._reuse is False:
            reuse = parent_reuse        # TF: reuse=False/None inherits the parent's flag
        STATE.scope.append((self._name, reuse))
        self.name = current_scope_name()
        return self

    def __exit__(self, *a):
        # TF (variable_scope.py, close_variable_subscopes): leaving a scope resets the default-name counters of
        # everything below it, so re-entering 'D_logit' with reuse=True names its tf.layers.dense 'dense' again
        full = self.name
        for k in [k for k in STATE.default_names if k == full or k.startswith(full + "/")]:
            del STATE.default_names[k]
        STATE.scope.pop()


AUTO_REUSE = "AUTO_REUSE"


def current_scope_name():
    return "/".join(n for n, _ in STATE.scope if n)


def _unique(table, parent, base):
    d = table.setdefault(parent, {})
    k = d.get(base, 0)
    d[base] = k + 1
    return base if k == 0 else "%s_%d" % (base, k)


def variable_scope(name_or_scope=None, default_name=None, reuse=None, custom_getter=None, **kw):
    """tf.variable_scope: nested names joined by '/', `reuse` inherited by inner scopes; with
    name_or_scope=None the default_name is made unique within the parent ('prelu', 'prelu_1', ...)."""
    if custom_getter is not None:
        raise NotImplementedError("custom_getter (EMA sampling path) is off the training step")
    if name_or_scope is None:
        name_or_scope = _unique(STATE.default_names, current_scope_name(), default_name)
    if isinstance(name_or_scope, _Scope):
        name_or_scope = name_or_scope._name
    return _Scope(name_or_scope, reuse)


class _Init:
    def __init__(self, kind, value=0.0, mean=0.0, stddev=1.0):
        self.kind, self.value, self.mean, self.stddev = kind, value, mean, stddev


def truncated_normal_initializer(mean=0.0, stddev=1.0, seed=None, dtype=None):
    assert mean == 0.0
    return _Init("trunc_normal", stddev=stddev)


def random_normal_initializer(mean=0.0, stddev=1.0, seed=None, dtype=None):
    assert mean == 0.0 and stddev == 1.0
    return _Init("normal")


def constant_initializer(value=0.0, dtype=None):
    return _Init("const", value=value)


def orthogonal_initializer(gain=1.0, seed=None, dtype=None):
    return _Init("orthogonal")


def get_variable(name, shape=None, dtype=None, initializer=None, regularizer=None, trainable=True, **kw):
    """tf.get_variable: full name = enclosing scopes + name; a second request is legal only under
    reuse=True (returns the same variable), a first request only outside it (AUTO_REUSE: both)."""
    if isinstance(shape, DType):           # get_variable(name, shape, dtype, ...) positional form
        shape, dtype = dtype, shape
    if isinstance(shape, (int, np.integer)):
        shape = [shape]
    shape = [int(s) for s in (shape if shape is not None else [])]
    scope = current_scope_name()
    full = scope + "/" + name if scope else name
    reuse = STATE.scope[-1][1] if STATE.scope else False
    if full in STATE.vars:
        if not reuse:
            raise ValueError("Variable %s already exists, disallowed. Did you mean to set reuse=True?" % full)
        v = STATE.vars[full]
        assert list(v.shape) == shape, (full, list(v.shape), shape)
        return v
    if reuse is True:
        raise ValueError("Variable %s does not exist, or was not created with tf.get_variable()." % full)
    if STATE.initial_values is not None and full in STATE.initial_values:
        arr = np.asarray(STATE.initial_values[full])
        assert list(arr.shape) == shape, (full, arr.shape, shape)
    else:
        init = initializer if initializer is not None else _Init("glorot")
        if init.kind not in ("trunc_normal", "normal", "const"):
            raise NotImplementedError("initializer %s for %s" % (init.kind, full))
        arr = STATE.init_fn(full, shape, init.kind, init.value, init.stddev)
    v = Variable(full, torch.as_tensor(np.asarray(arr), dtype=FLOAT).reshape(shape), trainable)
    STATE.vars[full] = v
    if regularizer is not None:
        STATE.regularizers.append((regularizer, v))
    return v


def trainable_variables():
    return [v for v in STATE.vars.values() if v.trainable]


def global_variables():
    return list(STATE.vars.values())


def assign(ref, value):
    return ref.assign(value)


@contextlib.contextmanager
def control_dependencies(ops_):
    yield


def get_collection(key):
    return []


class GraphKeys:
    UPDATE_OPS = "update_ops"
    LOCAL_VARIABLES = "local_variables"
    GLOBAL_VARIABLES = "variables"
    REGULARIZATION_LOSSES = "regularization_losses"


def get_regularization_losses():
    """tf.losses.get_regularization_losses(): one scalar per variable created with a regularizer
    (evaluated on the variable's current value)."""
    return [r(v) for r, v in STATE.regularizers]


def gradients(ys, xs):
    """tf.gradients(ys, xs): d(sum ys)/d xs; None for unconnected xs."""
    if not isinstance(ys, (list, tuple)):
        ys = [ys]
    single = not isinstance(xs, (list, tuple))
    xl = [xs] if single else list(xs)
    total = sum(_t(y).sum() for y in ys)
    g = torch.autograd.grad(total, [x.t for x in xl], allow_unused=True, retain_graph=True)
    out = [None if gi is None else Tensor(gi) for gi in g]
    return out


# ---------------------------------------------------------------------------------------------
# math
# ---------------------------------------------------------------------------------------------

def _axes(axis, nd):
    if axis is None:
        return tuple(range(nd))
    if isinstance(axis, (list, tuple)):
        return tuple(int(a) % nd for a in axis)
    return (int(axis) % nd,)


def reduce_mean(x, axis=None, keepdims=False, name=None):
    x = _t(x)
    return Tensor(x.mean(dim=_axes(axis, x.dim()), keepdim=keepdims))


def reduce_sum(x, axis=None, keepdims=False, name=None):
    x = _t(x)
    return Tensor(x.sum(dim=_axes(axis, x.dim()), keepdim=keepdims))


def matmul(a, b, transpose_a=False, transpose_b=False):
    a, b = _t(a), _t(b)
    if transpose_a:
        a = a.transpose(-1, -2)
    if transpose_b:
        b = b.transpose(-1, -2)
    return Tensor(a @ b)


def transpose(x, perm=None):
    x = _t(x)
    if perm is None:
        perm = list(range(x.dim()))[::-1]
    return Tensor(x.permute(*[int(p) for p in perm]))


def _shape_list(shape):
    if isinstance(shape, Tensor):
        return [int(v) for v in shape.t.reshape(-1)]
    return [int(_raw(s)) for s in shape]


def reshape(x, shape, name=None):
    return Tensor(_t(x).reshape(_shape_list(shape)))


def concat(values, axis, name=None):
    return Tensor(torch.cat([_t(v) for v in values], dim=int(axis)))


def split(value, num_or_size_splits, axis=0, name=None):
    x = _t(value)
    if isinstance(num_or_size_splits, int):
        n = num_or_size_splits
        assert x.shape[axis] % n == 0
        sizes = [x.shape[axis] // n] * n
    else:
        sizes = [int(s) for s in num_or_size_splits]
        assert sum(sizes) == x.shape[axis], (sizes, x.shape)
    return [Tensor(p) for p in torch.split(x, sizes, dim=axis)]


def stack(values, axis=0, name=None):
    return Tensor(torch.stack([_t(v) for v in values], dim=axis))


def expand_dims(x, axis, name=None):
    return Tensor(_t(x).unsqueeze(axis))


def shape(x, name=None):
    return Tensor(torch.as_tensor(list(_t(x).shape), dtype=torch.int64))


def cast(x, dtype, name=None):
    x = _t(x)
    if dtype.name == "int32" and x.is_floating_point():
        return Tensor(torch.trunc(x).to(torch.int64))          # tf.cast float->int truncates toward zero
    return Tensor(x.to(dtype.torch()))


def range_(limit, dtype=None, **kw):
    return Tensor(torch.arange(int(_raw(limit)), dtype=torch.int64 if dtype is None else dtype.torch()))


def ones(shape, dtype=float32, name=None):
    return Tensor(torch.ones(_shape_list(shape), dtype=dtype.torch()))


def zeros(shape, dtype=float32, name=None):
    return Tensor(torch.zeros(_shape_list(shape), dtype=dtype.torch()))


def ones_like(x, **kw):
    return Tensor(torch.ones_like(_t(x)))


def zeros_like(x, **kw):
    return Tensor(torch.zeros_like(_t(x)))


def eye(n, dtype=float32):
    return Tensor(torch.eye(int(n), dtype=dtype.torch()))


def constant(value, dtype=None, shape=None, name=None):
    t = _t(value)
    if dtype is not None:
        t = t.to(dtype.torch())
    return Tensor(t)


def subtract(a, b, name=None):
    return Tensor(_t(a) - _t(b))


def add_n(inputs, name=None):
    tot = _t(inputs[0])
    for v in inputs[1:]:
        tot = tot + _t(v)
    return Tensor(tot)


def maximum(a, b, name=None):
    a_ = _t(a)
    return Tensor(torch.maximum(a_, _t(b, like=a_).to(a_.dtype)))


def minimum(a, b, name=None):
    a_ = _t(a)
    return Tensor(torch.minimum(a_, _t(b, like=a_).to(a_.dtype)))


def clip_by_value(x, lo, hi, name=None):
    x = _t(x)
    lo_ = _t(lo, like=x).to(x.dtype)
    hi_ = _t(hi, like=x).to(x.dtype)
    return Tensor(torch.minimum(torch.maximum(x, lo_), hi_))


def sqrt(x, name=None): return Tensor(torch.sqrt(_t(x)))
def square(x, name=None): return Tensor(_t(x) * _t(x))
def tanh(x, name=None): return Tensor(torch.tanh(_t(x)))
def sigmoid(x, name=None): return Tensor(torch.sigmoid(_t(x)))
def abs_(x, name=None): return Tensor(torch.abs(_t(x)))
def stop_gradient(x, name=None): return Tensor(_t(x).detach())


def squared_difference(a, b, name=None):
    d = _t(a) - _t(b, like=_t(a))
    return Tensor(d * d)


def norm(x, ord="euclidean", axis=None, keepdims=False, name=None):
    assert ord in ("euclidean", 2)
    x = _t(x)
    return Tensor(torch.sqrt((x * x).sum(dim=_axes(axis, x.dim()), keepdim=keepdims)))


def meshgrid(*args, indexing="xy"):
    return [Tensor(g) for g in torch.meshgrid(*[_t(a) for a in args], indexing=indexing)]


def pad(x, paddings, mode="CONSTANT", name=None, constant_values=0):
    """tf.pad.  CONSTANT: zeros.  REFLECT: mirror WITHOUT repeating the edge sample
    (t = [1,2,3], paddings [[2,2]] -> [3,2,1,2,3,2,1]); requires pad < size."""
    x = _t(x)
    pads = [(int(_raw(a)), int(_raw(b))) for a, b in paddings]
    assert len(pads) == x.dim()
    mode = mode.upper()
    for ax, (a, b) in enumerate(pads):
        if a == 0 and b == 0:
            continue
        n = x.shape[ax]
        if mode == "REFLECT":
            assert a < n and b < n
            idx = np.pad(np.arange(n), (a, b), mode="reflect")
            x = x.index_select(ax, torch.as_tensor(idx, dtype=torch.int64))
        elif mode == "CONSTANT":
            shp_a = list(x.shape); shp_a[ax] = a
            shp_b = list(x.shape); shp_b[ax] = b
            x = torch.cat([x.new_full(shp_a, constant_values), x, x.new_full(shp_b, constant_values)], dim=ax)
        else:
            raise NotImplementedError(mode)
    return Tensor(x)


def gather_nd(params, indices, batch_dims=0, name=None):
    """tf.gather_nd with batch_dims=1 and index depth 1 (the only form DiffAugment_tf.py uses):
    out[b, i, ...] = params[b, indices[b, i, 0], ...]."""
    p, i = _t(params), _t(indices)
    assert batch_dims == 1 and i.dim() == 3 and i.shape[-1] == 1
    rows = []
    for b in range(p.shape[0]):
        rows.append(p[b].index_select(0, i[b, :, 0]))
    return Tensor(torch.stack(rows))


def scatter_nd(indices, updates, shape, name=None):
    """tf.scatter_nd: zeros(shape) with updates ADDED at the index tuples (duplicates accumulate)."""
    idx, upd = _t(indices), _t(updates)
    shp = _shape_list(shape)
    assert idx.shape[-1] == len(shp)
    out = torch.zeros(shp, dtype=upd.dtype)
    flat = idx.reshape(-1, len(shp))
    out.index_put_(tuple(flat[:, k] for k in range(len(shp))), upd.reshape(-1), accumulate=True)
    return Tensor(out)


def random_uniform(shape, minval=0, maxval=None, dtype=float32, seed=None, name=None):
    """tf.random.uniform: floats in [minval, maxval) (default [0,1)), ints in [minval, maxval).
    Values come from STATE.random_queue when the caller supplied them (randomness is an INPUT of the
    parity vectors), else from a seeded host generator; every draw is logged."""
    shp = _shape_list(shape)
    lo = _raw(minval)
    hi = _raw(maxval)
    lo = int(lo) if dtype.name == "int32" else float(lo)
    if hi is None:
        assert dtype.name != "int32"
        hi = 1.0
    hi = int(hi) if dtype.name == "int32" else float(hi)
    if STATE.random_queue:
        val = np.asarray(STATE.random_queue.pop(0)).reshape(shp)
        assert (val >= lo).all() and (val < hi).all(), ("supplied random value outside [%s, %s)" % (lo, hi))
    elif dtype.name == "int32":
        val = STATE.random_rng.integers(lo, hi, shp)
    else:
        val = lo + (hi - lo) * STATE.random_rng.random(shp)
    STATE.random_log.append((tuple(shp), lo, hi, dtype.name, np.array(val)))
    return Tensor(torch.as_tensor(val, dtype=dtype.torch()))


# ---------------------------------------------------------------------------------------------
# tf.nn
# ---------------------------------------------------------------------------------------------

def _same_pads(size, k, s, d=1):
    """TF 'SAME': out = ceil(size/s); total = max((out-1)*s + (k-1)*d + 1 - size, 0); the smaller half
    goes BEFORE (top/left), the extra sample after."""
    out = -(-size // s)
    total = max((out - 1) * s + (k - 1) * d + 1 - size, 0)
    return total // 2, total - total // 2


def _conv2d_raw(x, w, stride, padding, dilation=1):
    """x NHWC, w HWIO (cross-correlation, like TF)."""
    kh, kw = int(w.shape[0]), int(w.shape[1])
    if padding == "SAME":
        pt, pb = _same_pads(x.shape[1], kh, stride, dilation)
        pl, pr = _same_pads(x.shape[2], kw, stride, dilation)
        x = pad(Tensor(x), [[0, 0], [pt, pb], [pl, pr], [0, 0]]).t
    elif padding != "VALID":
        raise ValueError(padding)
    y = F.conv2d(x.permute(0, 3, 1, 2), w.permute(3, 2, 0, 1), stride=stride, dilation=dilation)
    return y.permute(0, 2, 3, 1)


def conv2d(input, filter, strides, padding, dilations=1, name=None):
    """tf.nn.conv2d, NHWC / HWIO, strides [1,s,s,1]."""
    x, w = _t(input), _t(filter)
    assert strides[0] == 1 and strides[3] == 1 and strides[1] == strides[2]
    d = dilations if isinstance(dilations, int) else dilations[1]
    assert x.shape[-1] == w.shape[2], (x.shape, w.shape)
    return Tensor(_conv2d_raw(x, w, int(strides[1]), padding, int(d)))


def conv2d_transpose(value, filter, output_shape, strides, padding="SAME", name=None):
    """tf.nn.conv2d_transpose(value, filter [kh,kw,out_ch,in_ch], output_shape, strides, padding) is
    DEFINED as the gradient of conv2d(input of `output_shape`, filter, strides, padding) with respect to
    that input, evaluated at output-gradient = value.  Computed exactly so (double-differentiable)."""
    x, w = _t(value), _t(filter)
    assert strides[0] == 1 and strides[3] == 1 and strides[1] == strides[2]
    out_shape = _shape_list(output_shape)
    assert w.shape[2] == out_shape[3] and w.shape[3] == x.shape[3], (w.shape, out_shape, x.shape)
    probe = torch.zeros(out_shape, dtype=x.dtype, requires_grad=True)
    with torch.enable_grad():
        y = _conv2d_raw(probe, w, int(strides[1]), padding)       # filter read as HWIO: I=out_ch, O=in_ch
        assert tuple(y.shape) == tuple(x.shape), ("conv2d_transpose: output_shape inconsistent", y.shape, x.shape)
        (g,) = torch.autograd.grad(y, probe, grad_outputs=x, create_graph=True)
    return Tensor(g)


def bias_add(value, bias, name=None):
    return Tensor(_t(value) + _t(bias))


def relu(x, name=None): return Tensor(torch.relu(_t(x)))


def leaky_relu(x, alpha=0.2, name=None):
    x = _t(x)
    return Tensor(torch.maximum(x, alpha * x))


def softmax(logits, axis=-1, name=None):
    return Tensor(torch.softmax(_t(logits), dim=axis))


def moments(x, axes, keep_dims=False, name=None):
    """tf.nn.moments: mean and BIASED variance over `axes`."""
    x = _t(x)
    ax = _axes(axes, x.dim())
    mean = x.mean(dim=ax, keepdim=True)
    var = ((x - mean) ** 2).mean(dim=ax, keepdim=keep_dims)
    return Tensor(mean if keep_dims else mean.reshape(var.shape)), Tensor(var)


def batch_normalization(x, mean, variance, offset, scale, variance_epsilon, name=None):
    """tf.nn.batch_normalization: inv = rsqrt(var + eps) * scale;  x * inv + (offset - mean * inv)."""
    x = _t(x)
    inv = torch.rsqrt(_t(variance) + variance_epsilon)
    if scale is not None:
        inv = inv * _t(scale)
    off = -_t(mean) * inv
    if offset is not None:
        off = _t(offset) + off
    return Tensor(x * inv + off)


def l2_normalize(x, axis=None, epsilon=1e-12, name=None, dim=None):
    """tf.nn.l2_normalize: x * rsqrt(max(sum(x^2, axis), epsilon))."""
    x = _t(x)
    if axis is None:
        axis = dim
    ss = (x * x).sum(dim=_axes(axis, x.dim()), keepdim=True)
    return Tensor(x * torch.rsqrt(torch.clamp(ss, min=epsilon)))


def l2_loss(x, name=None):
    x = _t(x)
    return Tensor((x * x).sum() / 2)


def sigmoid_cross_entropy_with_logits(labels=None, logits=None, name=None):
    """max(x, 0) - x * z + log(1 + exp(-|x|))."""
    x, z = _t(logits), _t(labels)
    return Tensor(torch.clamp(x, min=0) - x * z + torch.log1p(torch.exp(-torch.abs(x))))


def depth_to_space(x, block_size, name=None):
    x = _t(x)
    b, h, w, c = x.shape
    r = block_size
    x = x.reshape(b, h, w, r, r, c // (r * r)).permute(0, 1, 3, 2, 4, 5)
    return Tensor(x.reshape(b, h * r, w * r, c // (r * r)))


# ---------------------------------------------------------------------------------------------
# tf.layers / tf.image
# ---------------------------------------------------------------------------------------------

def layers_flatten(x, name=None):
    x = _t(x)
    return Tensor(x.reshape(x.shape[0], -1))


def _pool(x, pool_size, strides, padding, kind):
    x = _t(x)
    assert pool_size == 2 and strides == 2
    ph = _same_pads(x.shape[1], 2, 2) if padding.upper() == "SAME" else (0, 0)
    pw = _same_pads(x.shape[2], 2, 2) if padding.upper() == "SAME" else (0, 0)
    assert ph == (0, 0) and pw == (0, 0), "odd maps (SAME pooling pads) are off the hot path"
    b, h, w, c = x.shape
    win = x.reshape(b, h // 2, 2, w // 2, 2, c)
    if kind == "max":
        return Tensor(win.amax(dim=(2, 4)))
    return Tensor(win.mean(dim=(2, 4)))


def max_pooling2d(x, pool_size, strides, padding="valid", name=None):
    return _pool(x, pool_size, strides, padding, "max")


def average_pooling2d(x, pool_size, strides, padding="valid", name=None):
    return _pool(x, pool_size, strides, padding, "avg")


def layers_dense(inputs, units, kernel_initializer=None, kernel_regularizer=None, use_bias=True, name=None, **kw):
    """tf.layers.dense: variables '<layer>/kernel' [in, units] and '<layer>/bias' (zeros), layer name
    'dense', 'dense_1', ... unique within the enclosing scope."""
    x = _t(inputs)
    lname = name or _unique(STATE.default_names, current_scope_name(), "dense")
    with variable_scope(lname):
        w = get_variable("kernel", [x.shape[-1], units], initializer=kernel_initializer, regularizer=kernel_regularizer)
        y = x @ w.t
        if use_bias:
            b = get_variable("bias", [units], initializer=constant_initializer(0.0))
            y = y + b.t
    return Tensor(y)


def layers_batch_normalization(inputs, momentum=0.99, epsilon=1e-3, training=False, name=None, renorm=False, **kw):
    """tf.layers.batch_normalization(axis=-1, center, scale): variables gamma (1), beta (0), moving_mean (0),
    moving_variance (1) under scope `name`.  training=True: normalise with the batch mean / BIASED variance;
    the update ops (collected in UPDATE_OPS, run by the reference with every train op, utils.py:255) move
    moving_mean / moving_variance by (1 - momentum) with an assign_sub (two instantiations of one layer in a run
    -- D(real), D(fake) with --bn_in_d -- both apply their delta), the fused NHWC kernel feeding the UNBIASED
    (n/(n-1)) variance into the moving average."""
    if renorm:
        raise NotImplementedError("batch renorm is off the hot path")
    x = _t(inputs)
    c = x.shape[-1]
    with variable_scope(name or _unique(STATE.default_names, current_scope_name(), "batch_normalization")):
        gamma = get_variable("gamma", [c], initializer=constant_initializer(1.0))
        beta = get_variable("beta", [c], initializer=constant_initializer(0.0))
        mm = get_variable("moving_mean", [c], initializer=constant_initializer(0.0), trainable=False)
        mv = get_variable("moving_variance", [c], initializer=constant_initializer(1.0), trainable=False)
    if training:
        ax = tuple(range(x.dim() - 1))
        n = 1
        for a in ax:
            n *= x.shape[a]
        mean = x.mean(dim=ax)
        var = ((x - mean) ** 2).mean(dim=ax)
        # moving_averages.assign_moving_average: variable -= (variable - value) * (1 - momentum)
        mm.assign_sub((mm.t - mean) * (1 - momentum))
        mv.assign_sub((mv.t - var * (n / max(n - 1, 1))) * (1 - momentum))
    else:
        mean, var = mm.t, mv.t
    return batch_normalization(x, mean, var, beta, gamma, epsilon)


def resize_nearest_neighbor(images, size, name=None):
    x = _t(images)
    b, h, w, c = x.shape
    nh, nw = int(size[0]), int(size[1])
    assert nh % h == 0 and nw % w == 0
    return Tensor(x.repeat_interleave(nh // h, dim=1).repeat_interleave(nw // w, dim=2))


# ---------------------------------------------------------------------------------------------
# module tree + install
# ---------------------------------------------------------------------------------------------

def _module(name, **attrs):
    m = types.ModuleType(name)
    for k, v in attrs.items():
        setattr(m, k, v)
    return m


def _unavailable(what):
    def f(*a, **k):
        raise NotImplementedError("%s is outside the hot path covered by the eager shim" % what)
    return f


def build_module():
    tf = _module(
        "tensorflow",
        float32=float32, int32=int32, AUTO_REUSE=AUTO_REUSE, GraphKeys=GraphKeys, Variable=_unavailable("tf.Variable"),
        variable_scope=variable_scope, get_variable=get_variable, trainable_variables=trainable_variables,
        global_variables=global_variables, assign=assign, control_dependencies=control_dependencies,
        get_collection=get_collection, gradients=gradients, reset_default_graph=reset_default_graph,
        truncated_normal_initializer=truncated_normal_initializer, random_normal_initializer=random_normal_initializer,
        constant_initializer=constant_initializer, orthogonal_initializer=orthogonal_initializer,
        reduce_mean=reduce_mean, reduce_sum=reduce_sum, matmul=matmul, transpose=transpose, reshape=reshape,
        concat=concat, split=split, stack=stack, expand_dims=expand_dims, shape=shape, cast=cast, range=range_,
        ones=ones, zeros=zeros, ones_like=ones_like, zeros_like=zeros_like, eye=eye, constant=constant,
        subtract=subtract, add_n=add_n, maximum=maximum, minimum=minimum, clip_by_value=clip_by_value, sqrt=sqrt,
        square=square, tanh=tanh, stop_gradient=stop_gradient, squared_difference=squared_difference, norm=norm,
        meshgrid=meshgrid, pad=pad, gather_nd=gather_nd, scatter_nd=scatter_nd, random_uniform=random_uniform,
    )
    tf.random = _module("tensorflow.random", uniform=random_uniform,
                        normal=_unavailable("tf.random.normal"), truncated_normal=_unavailable("tf.random.truncated_normal"))
    tf.math = _module("tensorflow.math", sigmoid=sigmoid, abs=abs_)
    tf.nn = _module("tensorflow.nn", conv2d=conv2d, conv2d_transpose=conv2d_transpose, bias_add=bias_add, relu=relu,
                    leaky_relu=leaky_relu, softmax=softmax, moments=moments, batch_normalization=batch_normalization,
                    l2_normalize=l2_normalize, l2_loss=l2_loss, depth_to_space=depth_to_space,
                    sigmoid_cross_entropy_with_logits=sigmoid_cross_entropy_with_logits)
    tf.layers = _module("tensorflow.layers", flatten=layers_flatten, dense=layers_dense,
                        batch_normalization=layers_batch_normalization, max_pooling2d=max_pooling2d,
                        average_pooling2d=average_pooling2d, conv2d=_unavailable("tf.layers.conv2d"),
                        conv2d_transpose=_unavailable("tf.layers.conv2d_transpose"))
    tf.image = _module("tensorflow.image", resize_nearest_neighbor=resize_nearest_neighbor)
    tf.losses = _module("tensorflow.losses", get_regularization_losses=get_regularization_losses)
    tf.summary = _module("tensorflow.summary")
    tf.train = _module("tensorflow.train", AdamOptimizer=_unavailable("tf.train.AdamOptimizer"))
    contrib = _module("tensorflow.contrib")
    contrib.slim = _module("tensorflow.contrib.slim")
    contrib.data = _module("tensorflow.contrib.data", prefetch_to_device=None, shuffle_and_repeat=None, map_and_batch=None)
    contrib.opt = _module("tensorflow.contrib.opt", MovingAverageOptimizer=_unavailable("MovingAverageOptimizer"))
    contrib.layers = _module("tensorflow.contrib.layers", l2_regularizer=_unavailable("l2_regularizer"))
    tf.contrib = contrib
    return tf


def install(init_fn, initial_values=None):
    """Register the shim as `tensorflow` (+ the contrib sub-modules the reference imports) in sys.modules.
    init_fn(full_name, shape, kind, value, stddev) -> ndarray supplies deterministic initial values."""
    assert "tensorflow" not in sys.modules or getattr(sys.modules["tensorflow"], "__eager_shim__", False), \
        "a real TensorFlow is already imported"
    STATE.reset()
    STATE.init_fn = init_fn
    STATE.initial_values = initial_values
    tf = build_module()
    tf.__eager_shim__ = True
    sys.modules["tensorflow"] = tf
    for sub in ("random", "math", "nn", "layers", "image", "losses", "summary", "train", "contrib"):
        sys.modules["tensorflow." + sub] = getattr(tf, sub)
    for sub in ("slim", "data", "opt", "layers"):
        sys.modules["tensorflow.contrib." + sub] = getattr(tf.contrib, sub)
    return tf
