"""Case table of the reference-generated golden vectors (`tests/golden/ref_ops.npz`, `ref_diffaug.npz`,
`ref_net64.npz`).  Shared by the generator (`oracle/make_reference_golden.py`, which runs the reference's own
ops.py / DiffAugment_tf.py / BigGAN.py from /root/reference over the eager TF-1 shim) and by the tests, which
replay the same calls on the oracle (CPU) and on the CUDA path (GPU).

A case = one call `fn(*inputs, **kw, opt=OPT-with-overrides)` made at the root variable scope.  Inputs are
drawn from numpy default_rng(seed) in the order listed; shapes and seeds repeat the cases of
tests/test_ops_gpu.py so that the CUDA path is exercised on shapes it is already known to take.
"""
import numpy as np

# opt dict of the canonical flag set (BigGAN.py:248-274, :593-600; GANBase.py:39-51); "act" is resolved
# to the implementation's relu by the caller
OPT = {"is_training": True, "bn_in_d": False, "act": "relu", "upsampling_method": "deconv4", "g_conv": "deconv3",
       "downsampling_method": "strided_conv3", "self_attention_bias": True, "sn": True,
       "bn": {"type": "batch_norm", "momentum": 0.98}, "conv": {"padding_type": "reflect", "sn": True, "regularizer": None},
       "fc_regularizer": None}


def make_opt(act, overrides=None):
    """act: the implementation's relu, or a dict {"relu": fn, "prelu": fn, "lrelu": fn} when a case selects the activation."""
    import copy
    o = copy.deepcopy(OPT)
    for k, v in (overrides or {}).items():
        if isinstance(v, dict):
            o[k].update(v)
        else:
            o[k] = v
    o["act"] = act[o["act"]] if isinstance(act, dict) else act
    return o


def _c(name, fn, inputs, kw, seed, opt=None, scale=None, shift=None, relu_after=False, var_init=None):
    return {"name": name, "fn": fn, "inputs": inputs, "kw": kw, "seed": seed, "opt": opt or {},
            "scale": scale or {}, "shift": shift or {}, "relu_after": relu_after, "var_init": var_init or {}}


OP_CASES = [
    # conv: ops.py:49-113  (B, H, Cin, Cout, k, s, pad, bias, padtype rows of CONV_CASES)
    _c("conv_k3s1_reflect", "conv", [("x", (2, 8, 8, 16))], dict(channels=32, kernel=3, stride=1, pad=1, use_bias=False), 0),
    _c("conv_k3s2_reflect", "conv", [("x", (2, 8, 8, 16))], dict(channels=32, kernel=3, stride=2, pad=1, use_bias=False), 0),
    _c("conv_rgb_in_s2", "conv", [("x", (3, 16, 16, 3))], dict(channels=32, kernel=3, stride=2, pad=1, use_bias=False), 0),
    _c("conv_rgb_out", "conv", [("x", (2, 8, 8, 32))], dict(channels=3, kernel=3, stride=1, pad=1, use_bias=False), 0),
    _c("conv_1x1_bias", "conv", [("x", (2, 8, 8, 64))], dict(channels=8, kernel=1, stride=1, pad=0, use_bias=True), 0),
    _c("conv_odd_size_s2", "conv", [("x", (2, 9, 9, 8))], dict(channels=16, kernel=3, stride=2, pad=1, use_bias=True), 0),
    _c("conv_zero_pad_s1", "conv", [("x", (2, 8, 8, 16))], dict(channels=16, kernel=3, stride=1, pad=1, use_bias=True), 0,
       opt={"conv": {"padding_type": "zero"}}),
    _c("conv_zero_pad_s2", "conv", [("x", (2, 8, 8, 16))], dict(channels=16, kernel=3, stride=2, pad=1, use_bias=False), 0,
       opt={"conv": {"padding_type": "zero"}}),
    _c("conv_k4s2_pad1", "conv", [("x", (2, 6, 6, 8))], dict(channels=8, kernel=4, stride=2, pad=1, use_bias=False), 0),
    _c("conv_tc_shapes", "conv", [("x", (4, 16, 16, 96))], dict(channels=96, kernel=3, stride=2, pad=1, use_bias=False), 0),
    # deconv: ops.py:116-139  (B, H, Cin, Cout, k, s rows of DECONV_CASES; bias on)
    _c("deconv_k4s2", "deconv", [("x", (2, 4, 4, 32))], dict(channels=16, kernel=4, stride=2, use_bias=True), 1),
    _c("deconv_k3s1", "deconv", [("x", (2, 8, 8, 16))], dict(channels=16, kernel=3, stride=1, use_bias=True), 1),
    _c("deconv_k3s2", "deconv", [("x", (2, 4, 4, 16))], dict(channels=8, kernel=3, stride=2, use_bias=True), 1),
    _c("deconv_k6s2", "deconv", [("x", (1, 4, 4, 8))], dict(channels=8, kernel=6, stride=2, use_bias=True), 1),
    _c("deconv_k4s2_odd", "deconv", [("x", (2, 5, 5, 64))], dict(channels=64, kernel=4, stride=2, use_bias=True), 1),
    _c("deconv_k4s1", "deconv", [("x", (2, 4, 4, 16))], dict(channels=16, kernel=4, stride=1, use_bias=True), 1),
    # fully_connected: ops.py:148-175 (SN path and tf.layers.dense path)
    _c("fc_sn", "fully_connected", [("x", (6, 1, 1, 52))], dict(units=40, scope="fc"), 2),
    _c("fc_dense", "fully_connected", [("x", (6, 1, 1, 52))], dict(units=40, scope="fc"), 2, opt={"conv": {"sn": False}}),
    # condition_batch_norm (+ReLU, as the resblocks use it): ops.py:611-643, :253-254
    _c("cbn_relu_4x4x64", "cond_bn", [("x", (4, 4, 4, 64)), ("z", (4, 1, 1, 20))], {}, 4, scale={"x": 1.5}, shift={"x": 0.3},
       relu_after=True),
    _c("cbn_relu_8x8x32", "cond_bn", [("x", (3, 8, 8, 32)), ("z", (3, 1, 1, 20))], {}, 4, scale={"x": 1.5}, shift={"x": 0.3},
       relu_after=True),
    _c("cbn_inference", "cond_bn", [("x", (2, 4, 4, 16)), ("z", (2, 1, 1, 12))], {}, 6, opt={"is_training": False}),
    # batch_norm: ops.py:580-585
    _c("bn_train", "bn", [("x", (3, 8, 8, 32))], {}, 5, scale={"x": 2.0}, shift={"x": -0.5}),
    # self_attention_2: ops.py:467-492, gamma preset to 0.1 so that the branch contributes
    _c("sa2_8x8x32", "self_attention_2", [("x", (2, 8, 8, 32))], dict(channels=32), 8, var_init={"self_attention/gamma": [0.1]}),
    _c("sa2_16x16x64", "self_attention_2", [("x", (2, 16, 16, 64))], dict(channels=64), 8, var_init={"self_attention/gamma": [0.1]}),
    _c("sa_full_4x4x32", "self_attention", [("x", (2, 4, 4, 32))], dict(channels=32), 10, var_init={"self_attention/gamma": [0.1]}),
    # residual blocks: ops.py:250-266, :293-313, :187-198
    _c("resblock_up_condition", "resblock_up_condition", [("x", (2, 4, 4, 64)), ("z", (2, 1, 1, 24))],
       dict(channels=32, use_bias=False, scope="resblock_up_4"), 11),
    _c("resblock_down_rgb", "resblock_down", [("x", (2, 8, 8, 3))], dict(channels=64, use_bias=False, scope="resblock_down_1"), 12),
    _c("resblock_down_32", "resblock_down", [("x", (2, 8, 8, 32))], dict(channels=64, use_bias=False, scope="resblock_down_1"), 12),
    _c("resblock", "resblock", [("x", (2, 4, 4, 64))], dict(channels=64, use_bias=False), 13),
]

# SURVEY 8(f) "next" rows (ranks 1 and 3): activations and up/down-sampling / g_conv variants, bn_in_d.  -> ref_ops_next.npz
NEXT_OP_CASES = [
    _c("prelu", "prelu", [("x", (2, 4, 4, 16))], {}, 40,
       var_init={"prelu/alpha": [0.25, -0.5, 0.1, 0.0, 1.0, 0.3, -0.2, 0.05, 0.6, 0.15, -1.0, 0.2, 0.45, 0.0, 0.35, 0.8]}),
    _c("lrelu", "lrelu", [("x", (2, 4, 4, 16))], {}, 41),
    _c("upconv_deconv3", "upconv", [("x", (2, 4, 4, 16))], dict(channels=8), 42, opt={"upsampling_method": "deconv3"}),
    _c("upconv_deconv6", "upconv", [("x", (2, 4, 4, 16))], dict(channels=8), 42, opt={"upsampling_method": "deconv6"}),
    _c("upconv_subpixel2", "upconv", [("x", (2, 4, 4, 16))], dict(channels=8), 42, opt={"upsampling_method": "subpixel2"}),
    _c("upconv_subpixel3", "upconv", [("x", (2, 4, 4, 16))], dict(channels=8), 42, opt={"upsampling_method": "subpixel3"}),
    _c("upconv_resize_conv", "upconv", [("x", (2, 4, 4, 16))], dict(channels=8), 42, opt={"upsampling_method": "resize_conv"}),
    _c("upconv_nn", "upconv", [("x", (2, 4, 4, 16))], dict(channels=8), 42, opt={"upsampling_method": "nn"}),
    _c("g_conv_deconv4", "g_conv", [("x", (2, 8, 8, 16))], dict(channels=8), 43, opt={"g_conv": "deconv4"}),
    _c("g_conv_conv3", "g_conv", [("x", (2, 8, 8, 16))], dict(channels=8), 43, opt={"g_conv": "conv3"}),
    _c("g_conv_conv5", "g_conv", [("x", (2, 8, 8, 16))], dict(channels=8), 43, opt={"g_conv": "conv5"}),
    _c("downconv_resize_conv1", "downconv", [("x", (2, 8, 8, 16))], dict(channels=8), 44, opt={"downsampling_method": "resize_conv1"}),
    _c("downconv_resize_conv3", "downconv", [("x", (2, 8, 8, 16))], dict(channels=8), 44, opt={"downsampling_method": "resize_conv3"}),
    _c("downconv_pool_only", "downconv", [("x", (2, 8, 8, 16))], dict(channels=16), 44, opt={"downsampling_method": "pool_only"}),
    _c("downconv_max_pool_only", "downconv", [("x", (2, 8, 8, 16))], dict(channels=16), 44, opt={"downsampling_method": "max_pool_only"}),
    _c("resblock_down_resize_conv1", "resblock_down", [("x", (2, 8, 8, 16))], dict(channels=32, use_bias=False, scope="resblock_down_1"), 45,
       opt={"downsampling_method": "resize_conv1"}),
    _c("resblock_down_bn_in_d", "resblock_down", [("x", (3, 8, 8, 16))], dict(channels=32, use_bias=True, scope="resblock_down_1"), 46,
       opt={"bn_in_d": True}),
    _c("resblock_bn_in_d_prelu", "resblock", [("x", (3, 4, 4, 16))], dict(channels=16, use_bias=False), 47, opt={"bn_in_d": True, "act": "prelu"},
       var_init={"resblock/res1/prelu/alpha": [0.2] * 16}),
    _c("resblock_up_condition_prelu", "resblock_up_condition", [("x", (2, 4, 4, 32)), ("z", (2, 1, 1, 12))],
       dict(channels=16, use_bias=False, scope="resblock_up_4"), 48, opt={"act": "prelu"},
       var_init={"resblock_up_4/res1/prelu/alpha": [0.3] * 32, "resblock_up_4/res2/prelu/alpha": [-0.2] * 16}),
    # conditional batch renorm (ops.py:645-715): fresh state (weight 0: plain CBN with r = d = 1), a warmed-up state with
    # clipping active, the shared-statistics variant, inference
    _c("cbrn_fresh", "cond_bn", [("x", (4, 4, 4, 16)), ("z", (4, 1, 1, 12))], {}, 50, opt={"bn": {"type": "batch_renorm"}},
       scale={"x": 1.5}, shift={"x": 0.3}),
    _c("cbrn_warm", "cond_bn", [("x", (4, 4, 4, 16)), ("z", (4, 1, 1, 12))], {}, 50, scale={"x": 1.5}, shift={"x": 0.3},
       opt={"bn": {"type": "batch_renorm", "renorm_clipping": {"rmax": 1.2, "dmax": 0.3}, "renorm_momentum": 0.8}},
       var_init={"batch_renorm/renorm_weight": 0.7, "batch_renorm/renorm_mean": [0.1 * i - 0.5 for i in range(16)],
                 "batch_renorm/renorm_var": [0.5 + 0.2 * i for i in range(16)]}),
    _c("cbrn_shared", "cond_bn", [("x", (4, 4, 4, 16)), ("z", (4, 1, 1, 12))], {}, 50, scale={"x": 1.5}, shift={"x": 0.3},
       opt={"bn": {"type": "batch_renorm", "shared_renorm": True}},
       var_init={"batch_renorm/pop_mean": [0.05 * i for i in range(16)], "batch_renorm/pop_var": [1.0 + 0.1 * i for i in range(16)]}),
    _c("cbrn_inference", "cond_bn", [("x", (2, 4, 4, 16)), ("z", (2, 1, 1, 12))], {}, 51,
       opt={"is_training": False, "bn": {"type": "batch_renorm"}},
       var_init={"batch_renorm/pop_mean": [0.05 * i for i in range(16)], "batch_renorm/pop_var": [1.0 + 0.1 * i for i in range(16)]}),
    _c("resblock_up_condition_lrelu_subpixel3_conv5", "resblock_up_condition", [("x", (2, 4, 4, 32)), ("z", (2, 1, 1, 12))],
       dict(channels=16, use_bias=True, scope="resblock_up_4"), 49, opt={"act": "lrelu", "upsampling_method": "subpixel3", "g_conv": "conv5"}),
]

# names of inputs that stay fp32 on the CUDA side in bf16 mode (the CBN condition vector)
FP32_INPUTS = ("z",)

SN_CASES = [(27, 32), (864, 96), (1536, 1), (864, 3), (1030, 24)]      # spectral_norm on a [K, C] matrix, two runs

LOSS_FLOODS = (0.0, 0.1, 5.0)

DIFFAUG_CASES = [  # (name, policy, shape)
    ("color", "color", (4, 8, 8, 3)),
    ("translation", "translation", (5, 16, 16, 3)),
    ("cutout", "cutout", (5, 16, 16, 3)),
    ("cutout_odd", "cutout", (4, 9, 9, 3)),
    ("full", "color,translation,cutout", (4, 32, 32, 3)),
]

# whole-network cases: the reference's generator()/discriminator() (BigGAN.py:246-582, :591-715), canonical flags,
# one D phase and one G phase (losses and gradients, no optimizer)
NETS = {
    "net64": dict(img_size=64, ch=32, batch_size=4, n_labels=6, z_dim=128, sa_gamma_init=0.1),
    # BASELINE configs 1-4 as wiring cases (block tables, attention placement, z split, names) at narrow width
    "net128da": dict(img_size=128, ch=16, batch_size=2, n_labels=5, z_dim=120, sa_gamma_init=0.1, da_policy="color,translation,cutout"),
    "net256": dict(img_size=256, ch=8, batch_size=2, n_labels=4, z_dim=64, sa_gamma_init=0.1),
    "net512": dict(img_size=512, ch=8, batch_size=2, n_labels=4, z_dim=64, sa_gamma_init=0.1),
}
NET64 = NETS["net64"]

# the fork's CLI defaults on top of the hinge configuration: prelu, ortho_cosine regulariser, flood levels,
# DiffAugment 'full' (main.py:49-50, :63, :66-67, :80)
NEXT_NETS = {
    "net64cli": dict(img_size=64, ch=16, batch_size=4, n_labels=6, z_dim=128, sa_gamma_init=0.1, da_policy="color,translation,cutout",
                     activation="prelu", g_regularization="ortho_cosine", g_regularization_factor=1e-4, d_flood=0.1, g_flood=0.05),
    "net64variants": dict(img_size=64, ch=16, batch_size=4, n_labels=6, z_dim=128, sa_gamma_init=0.1, activation="lrelu",
                          upsampling_method="resize_conv", g_conv="conv3", downsampling_method="resize_conv1", bn_in_d=True,
                          g_regularization="ortho", g_regularization_factor=1e-4),
}


WHOLE_LIMIT = 4096      # golden arrays up to this size are stored whole; larger ones as a strided sample + sum + norm


def diffaug_inputs(shape):
    x = np.random.default_rng(30).uniform(-1, 1, shape).astype(np.float32)
    dy = np.random.default_rng(31).standard_normal(shape).astype(np.float32)
    return x, dy


def check(gold, key, value, tol, what=None, l2=False):
    """Compare `value` with the golden entry `key` (whole array, or strided sample + sum + L2 norm).
    tol is relative to max|golden| (l2=False) or a relative L2 error (l2=True, the bf16-mode metric)."""
    what = what or key
    v = np.asarray(value, dtype=np.float64)
    assert np.isfinite(v).all(), what + ": non-finite values"
    if key in gold.files:
        ref = np.asarray(gold[key], dtype=np.float64)
        assert v.shape == ref.shape or v.size == ref.size, (what, v.shape, ref.shape)
        v = v.reshape(ref.shape)
        store_eps = 0.0
    else:
        size, stride, total, l2norm = gold[key + "#stat"]
        assert v.size == int(size), (what, v.shape, tuple(gold[key + "#shape"]))
        flat = v.ravel()
        ref = np.asarray(gold[key + "#sub"], dtype=np.float64)
        nrm = float(np.sqrt((flat * flat).sum()))
        assert abs(nrm - l2norm) <= 4 * max(tol, 1e-12) * max(l2norm, 1e-30) + 1e-300, \
            "%s: L2 norm %.12g vs golden %.12g" % (what, nrm, l2norm)
        v = flat[::int(stride)]
        store_eps = 1e-7            # the sample is stored in float32
    if tol == 0.0:                  # bit-exact (index work): the float32 values themselves must be equal
        assert np.array_equal(v.astype(np.float32), ref.astype(np.float32)), what + ": not bit-exact"
        return
    scale = max(float(np.abs(ref).max()), 1e-6)
    if l2:
        err = float(np.linalg.norm(v - ref)) / max(float(np.linalg.norm(ref)), 1e-6 * np.sqrt(ref.size))
    else:
        err = float(np.abs(v - ref).max()) / scale
    assert err <= tol + store_eps, "%s: error %.3e > %.1e" % (what, err, tol)


def golden_absmax(gold, key):
    return float(np.abs(gold[key] if key in gold.files else gold[key + "#sub"]).max())


def case_inputs(case):
    rng = np.random.default_rng(case["seed"])
    out = {}
    for name, shape in case["inputs"]:
        a = rng.standard_normal(shape)
        a = a * case["scale"].get(name, 1.0) + case["shift"].get(name, 0.0)
        out[name] = a.astype(np.float32)
    return out


def case_dy(case, shape):
    return np.random.default_rng(case["seed"] + 1000).standard_normal(shape).astype(np.float32)
