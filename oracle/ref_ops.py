"""CPU ORACLE -- TEST INFRASTRUCTURE ONLY, NOT PRODUCT CODE.

A literal torch-CPU restatement (float64 or float32) of the reference's operator surface
(david-jk/BigGAN-Tensorflow ``ops.py`` and ``DiffAugment_tf.py``).  Only ``tests/``,
``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` / ``--impl reference`` legs may
import this package; the product package ``biggan_tensorflow_b200`` never does.

PARITY: pinned to the reference's own SOURCE.  `tests/golden/ref_*.npz` are produced by importing ops.py,
utils.py, DiffAugment_tf.py and BigGAN.py unmodified from /root/reference and executing them over
`oracle/tf1_eager.py`, an eager stand-in for the slice of the TF-1.x API they call
(`oracle/make_reference_golden.py`); this restatement reproduces those vectors to 2e-9 in float64
(`tests/test_reference_golden.py`: every op with gradients and state, spectral-norm `u` over two runs,
losses, regularisers, DiffAugment, and the 64x64 generator / discriminator with identical variable
names and shapes).  NOT pinned: TensorFlow's own kernels (TF 1.x cannot be installed in this image,
SURVEY.md section 8c) -- the stand-in's primitives follow TF's documented semantics and are checked
against the API docs' worked examples and brute-force loops (`tests/test_tf1_eager_shim.py`).  Also kept:
the analytic known-answer tests (`tests/test_oracle_kat.py`), an independent numpy restatement of the
index ops (`oracle/ref_diffaug_np.py`) and float64 finite-difference gradient checks.

Every function cites the reference file:line it follows.  Tensors are NHWC like the reference.
TF-1.x op semantics follow SURVEY.md appendix A.4.
"""
import zlib

import numpy as np
import torch
import torch.nn.functional as F


# --------------------------------------------------------------------------------------------
# Variable store: the tf.variable_scope / tf.get_variable contract (names of SURVEY.md A.3)
# --------------------------------------------------------------------------------------------

def name_seed(name):
    return zlib.crc32(name.encode("utf-8")) & 0xFFFFFFFF


def init_array(name, shape, kind, value=0.0):
    """Deterministic initialiser keyed by the variable's full TF name.

    kind: 'trunc_normal' = tf.truncated_normal_initializer(0, 0.02) (ops.py:13; redraw beyond
    2 sigma), 'normal' = tf.random_normal_initializer() (ops.py:722), 'const'.
    """
    shape = tuple(int(s) for s in shape)
    if kind == "const":
        return np.full(shape, value, dtype=np.float32)
    rng = np.random.Generator(np.random.PCG64(name_seed(name)))
    n = int(np.prod(shape)) if len(shape) else 1
    if kind == "normal":
        return rng.standard_normal(n).astype(np.float32).reshape(shape)
    if kind == "trunc_normal":
        x = rng.standard_normal(n)
        bad = np.abs(x) > 2.0
        while bad.any():
            x[bad] = rng.standard_normal(int(bad.sum()))
            bad = np.abs(x) > 2.0
        return (x * 0.02).astype(np.float32).reshape(shape)
    raise ValueError(kind)


class RefStore:
    """name -> torch tensor.  Creates on first use (deterministic init) or takes `params`."""

    def __init__(self, dtype=torch.float64, params=None, strict=False):
        self.dtype = dtype
        self.vars = {}
        self.trainable = {}
        self.scope = []
        self.strict = strict
        self.sn_cache = {}          # per update-phase cache of spectral_norm outputs
        self.pending_u = {}
        self.phase_snapshot = {}    # moving statistics as they were when the phase began (assign_sub semantics)
        self.regularizers = {}      # full variable name -> callable (tf REGULARIZATION_LOSSES, one entry per variable)
        self.default_names = {}     # parent scope -> {base: count}  (tf.variable_scope(None, default_name=...))
        if params:
            for k, v in params.items():
                self.vars[k] = torch.as_tensor(np.asarray(v), dtype=dtype).clone()

    # tf.variable_scope
    class _Scope:
        def __init__(self, store, name):
            self.store, self.name = store, name

        def __enter__(self):
            self.store.scope.append(self.name)
            return self

        def __exit__(self, *a):
            # TF resets the default-name counters of everything below a scope when the scope is left, so the
            # second instantiation of a network (reuse=True) names its 'prelu' scopes the same way again
            full = self.store.full_name()
            for k in [k for k in self.store.default_names if k == full or k.startswith(full + "/")]:
                del self.store.default_names[k]
            self.store.scope.pop()

    def variable_scope(self, name=None, default_name=None):
        if name is None:                                   # ops.py:533: variable_scope(name_or_scope=None, default_name="prelu")
            d = self.default_names.setdefault(self.full_name(), {})
            k = d.get(default_name, 0)
            d[default_name] = k + 1
            name = default_name if k == 0 else "%s_%d" % (default_name, k)
        return RefStore._Scope(self, name)

    def full_name(self, name=None):
        parts = [p for p in self.scope]
        if name is not None:
            parts.append(name)
        return "/".join(parts)

    def get_variable(self, name, shape, kind="trunc_normal", value=0.0, trainable=True, regularizer=None):
        full = self.full_name(name)
        if regularizer is not None:
            self.regularizers[full] = regularizer
        if full not in self.vars:
            if self.strict:
                raise KeyError(full)
            self.vars[full] = torch.as_tensor(init_array(full, shape, kind, value), dtype=self.dtype)
        t = self.vars[full]
        assert tuple(t.shape) == tuple(int(s) for s in shape), (full, tuple(t.shape), shape)
        self.trainable[full] = trainable
        if trainable and not t.requires_grad:
            t.requires_grad_(True)
        return t

    def assign(self, full, value):
        self.vars[full] = value.detach().clone()

    def begin_phase(self):
        """One sess.run: spectral_norm of a variable is evaluated once, `u` advances once."""
        self.sn_cache = {}
        self.phase_snapshot = {}

    def regularization_losses(self):
        """tf.losses.get_regularization_losses(): regularizer(variable) for every variable created with one."""
        return [fn(self.vars[n]) for n, fn in self.regularizers.items()]

    def trainables(self, substr):
        return {k: v for k, v in self.vars.items() if self.trainable.get(k, False) and substr in k}

    def export(self):
        return {k: v.detach().cpu().numpy().copy() for k, v in self.vars.items()}


# --------------------------------------------------------------------------------------------
# ops.py
# --------------------------------------------------------------------------------------------

# Optional storage-precision emulation (tests only): when set, every activation tensor the CUDA bf16 mode
# stores in bf16 is rounded (forward value and backward gradient) while the arithmetic stays in the store dtype.
# Used to show which error levels are inherent to bf16 storage (tests/test_bf16_conditioning.py).
ACT_ROUND = None


def _r(x):
    return ACT_ROUND(x) if ACT_ROUND is not None else x


def l2_normalize(x, eps=1e-12):
    # tf.nn.l2_normalize(axis=None): x * rsqrt(max(sum(x^2), eps))
    return x * torch.rsqrt(torch.clamp((x * x).sum(), min=eps))


def spectral_norm(store, w, iteration=1):
    """ops.py:718-747.  State semantics: evaluated once per update phase."""
    full_u = store.full_name("u")
    if full_u in store.sn_cache:
        return store.sn_cache[full_u]
    w_shape = list(w.shape)
    wm = w.reshape(-1, w_shape[-1])                                   # :720
    u = store.get_variable("u", [1, w_shape[-1]], kind="normal", trainable=False)  # :722
    u_hat = u
    v_hat = None
    for _ in range(iteration):
        v_ = u_hat @ wm.detach().t()                                  # :732
        v_hat = l2_normalize(v_)                                      # :733
        u_ = v_hat @ wm.detach()                                      # :735
        u_hat = l2_normalize(u_)                                      # :736
    u_hat = u_hat.detach()                                            # :738
    v_hat = v_hat.detach()                                            # :739
    sigma = (v_hat @ wm) @ u_hat.t()                                  # :741
    store.assign(full_u, u_hat)                                       # :743
    w_norm = (wm / sigma).reshape(w_shape)                            # :744-745
    store.sn_cache[full_u] = w_norm
    store.last_sigma = float(sigma.detach())
    return w_norm


def _nchw(x):
    return x.permute(0, 3, 1, 2)


def _nhwc(x):
    return x.permute(0, 2, 3, 1)


def tf_same_pad(size, k, s):
    out = -(-size // s)
    total = max((out - 1) * s + k - size, 0)
    return total // 2, total - total // 2


def conv(store, x, channels, opt, kernel=4, stride=2, pad=0, dilation=1, use_bias=True, scope="conv_0"):
    """ops.py:49-113 (int kernel only)."""
    with store.variable_scope(scope):
        pads = None
        zero_same = False
        if pad > 0:
            pad_type = opt.get("conv", {}).get("padding_type", "reflect")        # :66
            h = x.shape[1]
            if h % stride == 0:                                                 # :68-71
                pad = pad * 2
            else:
                pad = max(kernel - (h % stride), 0)
            pad_top = int(pad // 2)
            pad_bottom = int(pad - pad_top)
            pad_left = int(pad // 2)
            pad_right = int(pad - pad_left)
            if pad_type == "zero":
                zero_same = True                                                # :80
            elif pad_type == "reflect":
                pads = (pad_left, pad_right, pad_top, pad_bottom)               # :82
            else:
                raise ValueError("Unsupported padding type: " + str(pad_type))
        if not opt.get("conv", {}).get("sn", True):
            raise NotImplementedError("tf.layers.conv2d path (sn=False) is out of scope")
        reg = opt.get("conv", {}).get("regularizer", None) if "generator" in store.full_name() else None   # :87-92
        w = store.get_variable("kernel", [kernel, kernel, x.shape[-1], channels], regularizer=reg)
        wn = spectral_norm(store, w)                                            # :94
        xn = _nchw(x)
        if pads is not None:
            xn = F.pad(xn, pads, mode="reflect")
        if zero_same:
            pt, pb = tf_same_pad(x.shape[1], kernel, stride)
            pl, pr = tf_same_pad(x.shape[2], kernel, stride)
            xn = F.pad(xn, (pl, pr, pt, pb))
        y = F.conv2d(xn, wn.permute(3, 2, 0, 1), stride=stride, dilation=dilation)
        y = _nhwc(y)
        if use_bias:
            b = store.get_variable("bias", [channels], kind="const", value=0.0)  # :97
            y = y + b
        return _r(y)


def deconv(store, x, channels, opt, kernel=4, stride=2, padding="SAME", use_bias=True, scope="deconv_0"):
    """ops.py:116-139; tf.nn.conv2d_transpose SAME == adjoint of the SAME conv."""
    with store.variable_scope(scope):
        if padding != "SAME":
            raise NotImplementedError("VALID deconv is off the hot path")
        B, H, W, Cin = x.shape
        w = store.get_variable("kernel", [kernel, kernel, channels, Cin],
                               regularizer=opt.get("conv", {}).get("regularizer", None))   # :127
        wn = spectral_norm(store, w)                                            # :128
        # forward conv this is the adjoint of: input size H*stride, SAME pad
        pb_h, _ = tf_same_pad(H * stride, kernel, stride)
        pb_w, _ = tf_same_pad(W * stride, kernel, stride)
        assert pb_h == pb_w
        y = F.conv_transpose2d(_nchw(x), wn.permute(3, 2, 0, 1), stride=stride, padding=pb_h)
        # conv_transpose2d yields (H-1)*s - 2p + k; crop / zero-extend to H*s (output_shape, :121)
        Ho, Wo = H * stride, W * stride
        if y.shape[2] < Ho or y.shape[3] < Wo:
            y = F.pad(y, (0, max(Wo - y.shape[3], 0), 0, max(Ho - y.shape[2], 0)))
        y = y[:, :, :Ho, :Wo]
        y = _nhwc(y)
        if use_bias:
            b = store.get_variable("bias", [channels], kind="const", value=0.0)
            y = y + b
        return _r(y)


def fully_connected(store, x, units, opt, use_bias=True, lrmul=1.0, scope="fully_0"):
    """ops.py:148-175 (lrmul == 1 on the hot path)."""
    with store.variable_scope(scope):
        x = x.reshape(x.shape[0], -1)                                           # :150
        channels = x.shape[-1]
        if opt.get("conv", {}).get("sn", True):
            reg = opt.get("fc_regularizer") if "generator" in store.full_name() else None   # :155-158
            w = store.get_variable("kernel", [channels, units], regularizer=reg)
            if use_bias:
                b = store.get_variable("bias", [units], kind="const", value=0.0)
                return x @ spectral_norm(store, w) + b                          # :163
            return x @ spectral_norm(store, w)                                  # :165
        # tf.layers.dense (no SN): variables live under '<scope>/dense/{kernel,bias}'
        with store.variable_scope("dense"):                                     # :167-173
            reg = opt.get("fc_regularizer") if "generator" in store.full_name() else None
            w = store.get_variable("kernel", [channels, units], regularizer=reg)
            y = x @ w
            if use_bias:
                b = store.get_variable("bias", [units], kind="const", value=0.0)
                y = y + b
            return y


def relu(x):
    return torch.relu(x)                                                        # ops.py:529-530


def tanh(x):
    return torch.tanh(x)                                                        # ops.py:539-540


def lrelu(x, alpha=0.2):
    return torch.maximum(x, alpha * x)                                          # ops.py:525-526 tf.nn.leaky_relu


def make_prelu(store):
    """ops.py:532-537 bound to a store: alpha [C] (init 0) under '<scope>/prelu[_k]/alpha'."""
    def prelu(x, scope=None, init_val=0.0):
        with store.variable_scope(scope, default_name="prelu"):
            alphas = store.get_variable("alpha", [x.shape[-1]], kind="const", value=init_val)
            pos = torch.relu(x)
            neg = alphas * (x - torch.abs(x)) * 0.5
            return _r(pos + neg)
    return prelu


def avg_pooling(x):
    return _nhwc(F.avg_pool2d(_nchw(x), 2, 2))                                  # ops.py:512-514 (even maps)


def up_sample(x, scale_factor=2):
    """ops.py:516-519 tf.image.resize_nearest_neighbor to an integer multiple."""
    return x.repeat_interleave(scale_factor, dim=1).repeat_interleave(scale_factor, dim=2)


def depth_to_space(x, r):
    b, h, w, c = x.shape                                                        # tf.nn.depth_to_space, NHWC
    x = x.reshape(b, h, w, r, r, c // (r * r)).permute(0, 1, 3, 2, 4, 5)
    return x.reshape(b, h * r, w * r, c // (r * r))


def subpixel_conv(store, x, channels, opt, kernel=3, scale=2, use_bias=True, scope="subpixel_conv_0"):
    """ops.py:23-27."""
    r2 = scale * scale
    x = conv(store, x, channels * r2, opt=opt, kernel=kernel, stride=1, use_bias=use_bias, pad=(kernel - 1) / 2.0, scope=scope)
    return depth_to_space(x, scale)


def global_sum_pooling(x):
    return x.sum(dim=(1, 2))                                                    # ops.py:503-506


def max_pooling(x):
    return _nhwc(F.max_pool2d(_nchw(x), 2, 2))                                  # ops.py:508-510


def condition_batch_norm(store, x, z, opt={}, scope="batch_norm"):
    """ops.py:611-643."""
    with store.variable_scope(scope):
        c = x.shape[-1]
        decay = opt.get("bn", {}).get("momentum", 0.98)                         # :617
        epsilon = 1e-05
        test_mean = store.get_variable("pop_mean", [c], kind="const", value=0.0, trainable=False)
        test_var = store.get_variable("pop_var", [c], kind="const", value=1.0, trainable=False)
        beta = fully_connected(store, z, units=c, scope="beta", opt=opt)        # :623
        gamma = fully_connected(store, z, units=c, scope="gamma", opt=opt)      # :624
        beta = beta.reshape(-1, 1, 1, c)
        gamma = gamma.reshape(-1, 1, 1, c)
        if opt["is_training"]:
            batch_mean = x.mean(dim=(0, 1, 2))                                  # :630 tf.nn.moments
            batch_var = ((x - batch_mean) ** 2).mean(dim=(0, 1, 2))             # biased
            store.assign(store.full_name("pop_mean"), test_mean * decay + batch_mean * (1 - decay))
            store.assign(store.full_name("pop_var"), test_var * decay + batch_var * (1 - decay))
            mean, var = batch_mean, batch_var
        else:
            mean, var = test_mean, test_var
        inv = torch.rsqrt(var + epsilon) * gamma                                # tf.nn.batch_normalization
        return _r(x * inv + (beta - mean * inv))                                # :638


def normalize_renorm_clipping_params(renorm_clipping):
    """ops.py:587-597."""
    if "rmax" not in renorm_clipping:
        renorm_clipping["rmax"] = 1.5
    if "dmax" not in renorm_clipping:
        renorm_clipping["dmax"] = 0.5
    if "rmax" in renorm_clipping and "rmin" not in renorm_clipping:
        renorm_clipping["rmin"] = 1.0 / renorm_clipping["rmax"]
    return renorm_clipping


def condition_batch_renorm(store, x, z, opt={}, scope="batch_renorm"):
    """ops.py:645-715: conditional batch RE-normalisation (training-time running statistics with a fade-in weight,
    correction factors r, d clipped and treated as constants)."""
    with store.variable_scope(scope):
        c = x.shape[-1]
        renorm_clipping = normalize_renorm_clipping_params(dict(opt.get("bn", {}).get("renorm_clipping", {})))   # :651
        test_decay = opt.get("bn", {}).get("momentum", 0.98)
        renorm_decay = opt.get("bn", {}).get("renorm_momentum", 0.9)
        shared = opt.get("bn", {}).get("shared_renorm", False)
        renorm_fadein_decay = opt.get("bn", {}).get("renorm_fadein_decay", 0.9999)
        if not shared:
            test_decay = renorm_decay                                           # :658-659
        epsilon = 1e-05
        test_mean = store.get_variable("pop_mean", [c], kind="const", value=0.0, trainable=False)
        test_var = store.get_variable("pop_var", [c], kind="const", value=1.0, trainable=False)
        if not shared:
            renorm_mean = store.get_variable("renorm_mean", [c], kind="const", value=0.0, trainable=False)
            renorm_var = store.get_variable("renorm_var", [c], kind="const", value=1.0, trainable=False)
            renorm_weight = store.get_variable("renorm_weight", [], kind="const", value=0.0, trainable=False)
        else:
            renorm_mean, renorm_var, renorm_weight = test_mean, test_var, 1.0   # :671-674
        beta = fully_connected(store, z, units=c, scope="beta", opt=opt).reshape(-1, 1, 1, c)
        gamma = fully_connected(store, z, units=c, scope="gamma", opt=opt).reshape(-1, 1, 1, c)
        if not opt["is_training"]:
            inv = torch.rsqrt(test_var + epsilon) * gamma                       # :711-715
            return _r(x * inv + (beta - test_mean * inv))
        batch_mean = x.mean(dim=(0, 1, 2))
        batch_var = ((x - batch_mean) ** 2).mean(dim=(0, 1, 2))
        rmax, rmin, dmax = renorm_clipping["rmax"], renorm_clipping["rmin"], renorm_clipping["dmax"]
        sigma = torch.sqrt(batch_var + epsilon)
        renorm_sigma = torch.sqrt(renorm_var + epsilon)
        weighted_renorm_sigma = renorm_weight * renorm_sigma + (1 - renorm_weight) * sigma          # :690
        weighted_renorm_mean = renorm_weight * renorm_mean + (1 - renorm_weight) * batch_mean       # :691
        r = torch.clamp(sigma / weighted_renorm_sigma, rmin, rmax).detach()                         # :692
        d = torch.clamp((batch_mean - weighted_renorm_mean) / weighted_renorm_sigma, -dmax, dmax).detach()   # :693
        store.assign(store.full_name("pop_mean"), test_mean * test_decay + batch_mean * (1 - test_decay))
        store.assign(store.full_name("pop_var"), test_var * test_decay + batch_var * (1 - test_decay))
        if not shared:
            store.assign(store.full_name("renorm_mean"), renorm_mean * renorm_decay + batch_mean * (1 - renorm_decay))
            store.assign(store.full_name("renorm_var"), renorm_var * renorm_decay + batch_var * (1 - renorm_decay))
            store.assign(store.full_name("renorm_weight"),
                         torch.as_tensor(renorm_weight) * renorm_fadein_decay + 1.0 * (1 - renorm_fadein_decay))
        inv = torch.rsqrt(batch_var + epsilon) * (r * gamma)                    # :707 batch_normalization(x, mean, var, beta + d gamma, r gamma)
        return _r(x * inv + ((beta + d * gamma) - batch_mean * inv))


def cond_bn(store, x, z, opt={}, scope="batch_norm"):
    """ops.py:563-578."""
    type = opt.get("bn", {}).get("type", "bn")
    if type == "batch_norm_broken_renorm":
        type = "batch_norm"
        if scope == "batch_norm":
            scope = "batch_renorm"
    if type == "bn" or type == "batch_norm":
        return condition_batch_norm(store, x, z, opt=opt, scope=scope)
    elif type == "batch_renorm":
        if scope == "batch_norm":
            scope = "batch_renorm"
        return condition_batch_renorm(store, x, z, opt=opt, scope=scope)
    raise ValueError("Unknown BN type: " + str(type))


def batch_norm(store, x, opt={}, scope="batch_norm"):
    """ops.py:580-585: tf.layers.batch_normalization(momentum, eps 1e-5, training)."""
    with store.variable_scope(scope):
        c = x.shape[-1]
        momentum = opt.get("bn", {}).get("momentum", 0.98)
        gamma = store.get_variable("gamma", [c], kind="const", value=1.0)
        beta = store.get_variable("beta", [c], kind="const", value=0.0)
        mm = store.get_variable("moving_mean", [c], kind="const", value=0.0, trainable=False)
        mv = store.get_variable("moving_variance", [c], kind="const", value=1.0, trainable=False)
        if opt["is_training"]:
            n = x.shape[0] * x.shape[1] * x.shape[2]
            mean = x.mean(dim=(0, 1, 2))
            var = ((x - mean) ** 2).mean(dim=(0, 1, 2))
            # fused BN updates moving_variance with the Bessel-corrected variance.  The update is an
            # assign_sub of (value_at_phase_start - batch_stat) * (1 - momentum): with one instantiation per phase this
            # is the usual moving average; two instantiations (D(real), D(fake) with bn_in_d) both apply their delta
            for nm, cur, stat in (("moving_mean", mm, mean), ("moving_variance", mv, var * (n / max(n - 1, 1)))):
                full = store.full_name(nm)
                snap = store.phase_snapshot.setdefault(full, cur.detach().clone())
                store.assign(full, cur.detach() - (snap - stat.detach()) * (1 - momentum))
        else:
            mean, var = mm, mv
        inv = torch.rsqrt(var + 1e-05) * gamma
        return _r(x * inv + (beta - mean * inv))


def bn(store, x, opt={}, scope="batch_norm"):
    """ops.py:546-561."""
    type = opt.get("bn", {}).get("type", "bn")
    if type == "bn" or type == "batch_norm":
        return batch_norm(store, x, opt=opt, scope=scope)
    raise ValueError("Unknown BN type: " + str(type))


def hw_flatten(x):
    return x.reshape(x.shape[0], -1, x.shape[-1])                               # ops.py:180-181


def self_attention(store, x, channels, opt, scope="self_attention"):
    """ops.py:445-465 (dead code in the reference; kept as API surface)."""
    with store.variable_scope(scope):
        use_bias = opt.get("self_attention_bias", False)
        f = conv(store, x, channels // 8, kernel=1, stride=1, opt=opt, scope="f_conv", use_bias=use_bias)
        g = conv(store, x, channels // 8, kernel=1, stride=1, opt=opt, scope="g_conv", use_bias=use_bias)
        h = conv(store, x, channels, kernel=1, stride=1, opt=opt, scope="h_conv", use_bias=use_bias)
        s = hw_flatten(g) @ hw_flatten(f).transpose(1, 2)
        beta = torch.softmax(s, dim=-1)
        o = beta @ hw_flatten(h)
        gamma = store.get_variable("gamma", [1], kind="const", value=0.0)
        o = o.reshape(x.shape)
        return gamma * o + x


def self_attention_2(store, x, channels, opt, scope="self_attention"):
    """ops.py:467-492 (the live attention block)."""
    with store.variable_scope(scope):
        use_bias = opt.get("self_attention_bias", False)                        # :470
        f = conv(store, x, channels // 8, kernel=1, stride=1, opt=opt, scope="f_conv", use_bias=use_bias)
        f = max_pooling(f)                                                      # :473
        g = conv(store, x, channels // 8, kernel=1, stride=1, opt=opt, scope="g_conv", use_bias=use_bias)
        h = conv(store, x, channels // 2, kernel=1, stride=1, opt=opt, scope="h_conv", use_bias=use_bias)
        h = max_pooling(h)                                                      # :478
        s = hw_flatten(g) @ hw_flatten(f).transpose(1, 2)                       # :481  [bs, N, N/4]
        beta = torch.softmax(s, dim=-1)                                         # :483
        o = _r(beta @ hw_flatten(h))                                            # :485
        gamma = store.get_variable("gamma", [1], kind="const", value=0.0)       # :486
        o = o.reshape(x.shape[0], x.shape[1], x.shape[2], channels // 2)        # :488
        o = conv(store, o, channels, kernel=1, stride=1, opt=opt, scope="attn_conv", use_bias=use_bias)
        return _r(gamma * o + x)                                                # :490


def upconv(store, x, channels, opt, use_bias=True):
    """ops.py:200-218."""
    m = opt["upsampling_method"]
    if m == "deconv3":
        return deconv(store, x, channels, kernel=3, stride=2, use_bias=use_bias, opt=opt)
    if m == "deconv4":
        return deconv(store, x, channels, kernel=4, stride=2, use_bias=use_bias, opt=opt)
    if m == "deconv6":
        return deconv(store, x, channels, kernel=6, stride=2, use_bias=use_bias, opt=opt)
    if m == "subpixel2":
        return subpixel_conv(store, x, channels, kernel=2, scale=2, use_bias=use_bias, opt=opt)
    if m == "subpixel3":
        return subpixel_conv(store, x, channels, kernel=3, scale=2, use_bias=use_bias, opt=opt)
    if m == "resize_conv":
        x = up_sample(x, 2)
        return conv(store, x, channels, kernel=3, stride=1, pad=1, use_bias=use_bias, opt=opt)
    if m == "nn":
        return up_sample(x, 2)
    raise ValueError("Invalid upsampling method specified: " + str(m))


def g_conv(store, x, channels, opt, use_bias=True):
    """ops.py:220-230."""
    m = opt["g_conv"]
    if m == "deconv3":
        return deconv(store, x, channels, kernel=3, stride=1, use_bias=use_bias, opt=opt)
    if m == "deconv4":
        return deconv(store, x, channels, kernel=4, stride=1, use_bias=use_bias, opt=opt)
    if m == "conv3":
        return conv(store, x, channels, kernel=3, stride=1, pad=1, use_bias=use_bias, opt=opt)
    if m == "conv5":
        return conv(store, x, channels, kernel=5, stride=1, pad=2, use_bias=use_bias, opt=opt)
    raise ValueError("Invalid generator convolution type specified: " + str(m))


def downconv(store, x, channels, opt, use_bias=True, method=None):
    """ops.py:269-291."""
    if method is None:
        method = opt["downsampling_method"]
    if method == "strided_conv3":
        return conv(store, x, channels, kernel=3, stride=2, pad=1, use_bias=use_bias, opt=opt)
    if method == "resize_conv1":
        x = conv(store, x, channels, kernel=1, stride=1, pad=0, use_bias=use_bias, opt=opt)
        return avg_pooling(x)
    if method == "resize_conv3":
        x = conv(store, x, channels, kernel=3, stride=1, pad=1, use_bias=use_bias, opt=opt)
        return avg_pooling(x)
    if method == "pool_only":
        return avg_pooling(x)
    if method == "max_pool_only":
        return max_pooling(x)
    raise ValueError("Invalid downsampling method specified: " + str(method))


def resblock(store, x_init, channels, opt, use_bias=True, scope="resblock"):
    """ops.py:187-198."""
    with store.variable_scope(scope):
        with store.variable_scope("res1"):
            x = conv(store, x_init, channels, kernel=3, stride=1, pad=1, use_bias=use_bias, opt=opt)
            if opt["bn_in_d"]:
                x = bn(store, x, opt=opt)
            x = opt["act"](x)
        with store.variable_scope("res2"):
            x = conv(store, x, channels, kernel=3, stride=1, pad=1, use_bias=use_bias, opt=opt)
            if opt["bn_in_d"]:
                x = bn(store, x, opt=opt)
        return x + x_init


def resblock_up_condition(store, x_init, z, channels, opt, use_bias=True, scope="resblock_up"):
    """ops.py:250-266."""
    with store.variable_scope(scope):
        with store.variable_scope("res1"):
            x = cond_bn(store, x_init, z, opt=opt)
            x = opt["act"](x)
            x = upconv(store, x, channels, use_bias=use_bias, opt=opt)
        with store.variable_scope("res2"):
            x = cond_bn(store, x, z, opt=opt)
            x = opt["act"](x)
            x = g_conv(store, x, channels, use_bias=use_bias, opt=opt)
        with store.variable_scope("skip"):
            x_init = upconv(store, x_init, channels, use_bias=use_bias, opt=opt)
    return _r(x + x_init)


def resblock_down(store, x_init, channels, opt, use_bias=True, scope="resblock_down"):
    """ops.py:293-313."""
    with store.variable_scope(scope):
        with store.variable_scope("res1"):
            if opt["bn_in_d"]:
                x = bn(store, x_init, opt=opt)
            else:
                x = x_init
            x = opt["act"](x)
            res_method = opt["downsampling_method"]
            if res_method != "strided_conv3":
                res_method = "resize_conv3"
            x = downconv(store, x, channels, use_bias=use_bias, opt=opt, method=res_method)
        with store.variable_scope("res2"):
            if opt["bn_in_d"]:
                x = bn(store, x, opt=opt)
            x = opt["act"](x)
            x = conv(store, x, channels, kernel=3, stride=1, pad=1, use_bias=use_bias, opt=opt)
        with store.variable_scope("skip"):
            x_init = downconv(store, x_init, channels, use_bias=use_bias, opt=opt,
                              method=opt["downsampling_method"])
    return _r(x + x_init)


def flood_loss(loss, flood_level):
    return torch.abs(loss - flood_level) + flood_level                          # ops.py:847-848


def discriminator_loss(loss_func, real, fake, flood_level=0):
    """ops.py:753-797 ('hinge' only)."""
    if loss_func != "hinge":
        raise NotImplementedError(loss_func)
    real_loss = torch.relu(1.0 - real).mean()                                   # :789
    fake_loss = torch.relu(1.0 + fake).mean()                                   # :790
    loss = real_loss + fake_loss
    if flood_level:
        loss = flood_loss(loss, flood_level)
    return loss


def generator_loss(loss_func, fake, real, flood_level=0):
    """ops.py:799-840 ('hinge' only)."""
    if loss_func != "hinge":
        raise NotImplementedError(loss_func)
    loss = -fake.mean()                                                         # :833
    if flood_level:
        loss = flood_loss(loss, flood_level)
    return loss


def cls_loss_logistic(truth, answer, cls_weights=1.0):
    """utils.py:366-369: mean(sigmoid_cross_entropy_with_logits(labels, logits) * w)."""
    xent = torch.clamp(answer, min=0) - answer * truth + torch.log1p(torch.exp(-torch.abs(answer)))
    return (xent * cls_weights).mean()


def cosine_similarity(a, b):
    """utils.py:180-183: rows l2-normalised (tf.nn.l2_normalize(axis=1), eps 1e-12), then a b^T."""
    na = a * torch.rsqrt(torch.clamp((a * a).sum(dim=1, keepdim=True), min=1e-12))
    nb = b * torch.rsqrt(torch.clamp((b * b).sum(dim=1, keepdim=True), min=1e-12))
    return na @ nb.t()


def orthogonal_regularizer(scale, type="ortho"):
    """utils.py:185-235 (conv and FC forms differ only in the reshape): scale * l2_loss(reg),
    reg = W^T W - I ('ortho') or cosine_similarity(W^T W, 1 - I) ('ortho_cosine')."""
    def ortho_reg(w):
        c = w.shape[-1]
        wm = w.reshape(-1, c)                                                   # :192
        identity = torch.eye(c, dtype=w.dtype)                                  # :195
        w_mul = wm.t() @ wm                                                     # :197-198
        if type == "ortho":
            reg = w_mul - identity                                              # :200
        elif type == "ortho_cosine":
            reg = cosine_similarity(w_mul, torch.ones_like(identity) - identity)   # :202
        else:
            raise ValueError("Unknown regularization method.")
        return scale * (reg * reg).sum() / 2                                    # :207-209 tf.nn.l2_loss
    return ortho_reg


orthogonal_regularizer_fc = orthogonal_regularizer


# --------------------------------------------------------------------------------------------
# DiffAugment_tf.py -- randomness is an INPUT (dict of per-sample arrays), see SURVEY 8d
# --------------------------------------------------------------------------------------------

def translation_shift(size, ratio=0.125):
    return int(size * ratio + 0.5)                                              # DiffAugment_tf.py:43


def cutout_size(size, ratio=0.5):
    return int(size * ratio + 0.5)                                              # DiffAugment_tf.py:56


def rand_brightness(x, rnd):
    magnitude = rnd["brightness"].reshape(-1, 1, 1, 1).to(x.dtype) - 0.5        # :21
    return x + magnitude


def rand_saturation(x, rnd):
    magnitude = rnd["saturation"].reshape(-1, 1, 1, 1).to(x.dtype) * 2          # :27
    x_mean = x.mean(dim=3, keepdim=True)
    return (x - x_mean) * magnitude + x_mean


def rand_contrast(x, rnd):
    magnitude = rnd["contrast"].reshape(-1, 1, 1, 1).to(x.dtype) + 0.5          # :34
    x_mean = x.mean(dim=(1, 2, 3), keepdim=True)
    return (x - x_mean) * magnitude + x_mean


def rand_translation(x, rnd, ratio=0.125):
    """DiffAugment_tf.py:40-50 with translation_x / translation_y given (int, per sample)."""
    B, H, W, C = x.shape
    tx = rnd["translation_x"].reshape(B, 1).long()
    ty = rnd["translation_y"].reshape(B, 1).long()
    grid_x = torch.clamp(torch.arange(H).reshape(1, H) + tx + 1, 0, H + 1)      # :46
    grid_y = torch.clamp(torch.arange(W).reshape(1, W) + ty + 1, 0, W + 1)      # :47
    xp = F.pad(x, (0, 0, 0, 0, 1, 1))                                           # pad axis 1 (H)
    x = torch.gather(xp, 1, grid_x.reshape(B, H, 1, 1).expand(B, H, W, C))      # :48
    xp = F.pad(x, (0, 0, 1, 1))                                                 # pad axis 2 (W)
    x = torch.gather(xp, 2, grid_y.reshape(B, 1, W, 1).expand(B, H, W, C))      # :49
    return x


def rand_cutout(x, rnd, ratio=0.5):
    """DiffAugment_tf.py:53-66 with offset_x / offset_y given."""
    B, H, W, C = x.shape
    cs0, cs1 = cutout_size(H, ratio), cutout_size(W, ratio)
    ox = rnd["cutout_x"].reshape(B, 1, 1).long()
    oy = rnd["cutout_y"].reshape(B, 1, 1).long()
    gb, gx, gy = torch.meshgrid(torch.arange(B), torch.arange(cs0), torch.arange(cs1), indexing="ij")
    gx = torch.clamp(gx + ox - cs0 // 2, 0, H - 1)                              # :61-63
    gy = torch.clamp(gy + oy - cs1 // 2, 0, W - 1)
    mask = torch.ones(B, H, W, dtype=x.dtype)
    mask[gb.reshape(-1), gx.reshape(-1), gy.reshape(-1)] = 0                    # :64 max(1-scatter,0)
    return x * mask.unsqueeze(3)                                                # :65


AUGMENT_FNS = {                                                                 # :69-73
    "color": [rand_brightness, rand_saturation, rand_contrast],
    "translation": [rand_translation],
    "cutout": [rand_cutout],
}


def DiffAugment(x, policy="", channels_first=False, rnd=None):
    """DiffAugment_tf.py:8-17."""
    if policy:
        if channels_first:
            x = x.permute(0, 2, 3, 1)
        for p in policy.split(","):
            for f in AUGMENT_FNS[p]:
                x = f(x, rnd)
        if channels_first:
            x = x.permute(0, 3, 1, 2)
    return x


def draw_diffaug_randoms(rng, B, H, W):
    """Host-side draws in the ranges DiffAugment_tf.py uses (tf.random.uniform [min,max))."""
    sh, sw = translation_shift(H), translation_shift(W)
    cs0, cs1 = cutout_size(H), cutout_size(W)
    return {
        "brightness": rng.random(B, dtype=np.float32),
        "saturation": rng.random(B, dtype=np.float32),
        "contrast": rng.random(B, dtype=np.float32),
        "translation_x": rng.integers(-sh, sh + 1, B).astype(np.int32),
        "translation_y": rng.integers(-sw, sw + 1, B).astype(np.int32),
        "cutout_x": rng.integers(0, H + (1 - cs0 % 2), B).astype(np.int32),
        "cutout_y": rng.integers(0, W + (1 - cs1 % 2), B).astype(np.int32),
    }


def rnd_to_torch(rnd):
    return {k: torch.as_tensor(v) for k, v in rnd.items()}
