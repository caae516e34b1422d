"""GPU parity tests: every operator of the hot path (CUDA, called through the C ABI) against the CPU
oracle (float64 restatement of the reference) on the same seeded inputs and identical variables.
Tolerances are north_star's: 1e-4 relative in fp32 mode, 2e-2 in bf16 mode (relative to max|ref|).
"""
import numpy as np
import pytest
import torch

from helpers import TOL, assert_close, dev_tensor, new_product_store, seed_backward
from oracle import ref_ops as R

pytestmark = pytest.mark.gpu

DTYPES = [torch.float32, torch.bfloat16]


def _opt(is_training=True, pad="reflect"):
    from biggan_tensorflow_b200 import ops
    return ({"is_training": is_training, "bn_in_d": False, "act": ops.relu, "upsampling_method": "deconv4",
             "g_conv": "deconv3", "downsampling_method": "strided_conv3", "self_attention_bias": True,
             "bn": {"type": "batch_norm", "momentum": 0.98}, "conv": {"padding_type": pad, "sn": True},
             "fc_regularizer": None},
            {"is_training": is_training, "bn_in_d": False, "act": R.relu, "upsampling_method": "deconv4",
             "g_conv": "deconv3", "downsampling_method": "strided_conv3", "self_attention_bias": True,
             "bn": {"type": "batch_norm", "momentum": 0.98}, "conv": {"padding_type": pad, "sn": True},
             "fc_regularizer": None})


def run_case(prod_fn, ref_fn, inputs, dtype, tol=None, fp32_inputs=(), check_state=True, seed=0, tol_scale=1.0):
    """inputs: name -> numpy array (differentiable activations).  prod_fn(xs) / ref_fn(store, xs)."""
    tol = (tol or TOL[dtype]) * tol_scale
    rng = np.random.default_rng(seed + 1000)
    st = new_product_store(dtype)
    xs = {k: dev_tensor(v, torch.float32 if k in fp32_inputs else dtype) for k, v in inputs.items()}
    y = prod_fn(xs)
    dy = rng.standard_normal(y.shape).astype(np.float32)
    seed_backward(y, dy)

    rs = R.RefStore(torch.float64)
    xr = {k: torch.tensor(v, dtype=torch.float64, requires_grad=True) for k, v in inputs.items()}
    yr = ref_fn(rs, xr)
    tnames = [n for n in rs.vars if rs.trainable.get(n, False)]
    wrt = list(xr.values()) + [rs.vars[n] for n in tnames]
    grads = torch.autograd.grad((yr * torch.as_tensor(dy, dtype=torch.float64)).sum(), wrt, allow_unused=True)
    assert tuple(y.shape) == tuple(yr.shape)
    assert_close(y.numpy(), yr.detach().numpy(), tol, "forward")
    for (k, _), g in zip(xr.items(), grads[:len(xr)]):
        if g is None:
            continue
        assert xs[k].grad is not None, "no gradient reached input " + k
        assert_close(xs[k].grad.float().cpu().numpy(), g.numpy(), tol, "d/d" + k)
    assert set(st.vars.keys()) == set(rs.vars.keys()), (sorted(set(st.vars) ^ set(rs.vars)))
    gscale = max([float(g.abs().max()) for g in grads[len(xr):] if g is not None] + [0.0])
    for n, g in zip(tnames, grads[len(xr):]):
        if g is None:
            continue
        pv = st.vars[n]
        assert pv.grad is not None, "no gradient reached variable " + n
        if n.endswith("f_conv/bias"):
            # softmax is invariant to a shift of all keys: this gradient is analytically zero and both
            # sides only hold summation noise
            assert float(g.abs().max()) <= 1e-9 * max(gscale, 1.0)
            assert float(pv.grad.float().abs().max()) <= 2e-3 * gscale, n      # bf16 P / dS rounding noise, relative to the largest gradient
            continue
        # max-pooled attention branches under bf16: see tests/test_bf16_conditioning.py
        vtol = 1.5e-1 if (tol >= 1e-3 and ("f_conv/kernel" in n or "h_conv/kernel" in n)) else tol
        if tol >= 1e-3 and n.endswith("g_conv/bias"):
            vtol = 1.5e-1          # sum over all queries of dQ (mixed signs): bf16 P / dS operand rounding does not average out
        assert_close(pv.grad.float().cpu().numpy(), g.numpy(), vtol, "d/d" + n)
    if check_state:
        for n, v in rs.vars.items():
            if not rs.trainable.get(n, False):
                assert_close(st.vars[n].data.float().cpu().numpy(), v.detach().numpy(), max(TOL[dtype], 1e-4), "state " + n)
    return st, rs


# ------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("K,C", [(27, 32), (864, 96), (1536, 1), (864, 3), (1030, 24), (4096, 1536), (300, 4100)])
def test_spectral_norm(K, C):
    """ops.py:718-747: W/sigma, u update, gradient with u_hat / v_hat frozen."""
    from biggan_tensorflow_b200 import ops
    from biggan_tensorflow_b200.tf_shim import get_variable, variable_scope

    def prod(xs):
        with variable_scope("t"):
            w = get_variable("kernel", [K, C], initializer=ops.weight_init)
            return ops.spectral_norm(w)

    def ref(rs, xs):
        with rs.variable_scope("t"):
            w = rs.get_variable("kernel", [K, C])
            return R.spectral_norm(rs, w)

    run_case(prod, ref, {}, torch.float32)


def test_spectral_norm_rank1_known_answer():
    """KAT: for W = a b^T one power iteration gives sigma = |a||b| exactly."""
    from biggan_tensorflow_b200 import ops
    from biggan_tensorflow_b200.tf_shim import get_variable, variable_scope
    st = new_product_store(torch.float32)
    rng = np.random.default_rng(3)
    a, b = rng.standard_normal(40), rng.standard_normal(24)
    with variable_scope("t"):
        w = get_variable("kernel", [40, 24], initializer=ops.weight_init)
        w.data.copy_(torch.from_numpy(np.outer(a, b).astype(np.float32)))
        wn = ops.spectral_norm(w)
    sigma = float(wn.aux["sigma"].cpu())
    assert abs(sigma - np.linalg.norm(a) * np.linalg.norm(b)) < 1e-4 * sigma
    s = np.linalg.svd(wn.numpy(), compute_uv=False)
    assert abs(s[0] - 1.0) < 1e-4


CONV_CASES = [
    # B, H, Cin, Cout, k, s, pad, bias, padtype
    (2, 8, 16, 32, 3, 1, 1, False, "reflect"),
    (2, 8, 16, 32, 3, 2, 1, False, "reflect"),
    (3, 16, 3, 32, 3, 2, 1, False, "reflect"),      # first D conv (Cin = 3)
    (2, 8, 32, 3, 3, 1, 1, False, "reflect"),       # G_logit (Cout = 3)
    (2, 8, 64, 8, 1, 1, 0, True, "reflect"),        # attention 1x1 with bias
    (2, 9, 8, 16, 3, 2, 1, True, "reflect"),        # odd size: pad = max(k - h % s, 0)
    (2, 8, 16, 16, 3, 1, 1, True, "zero"),
    (2, 8, 16, 16, 3, 2, 1, False, "zero"),
    (1, 4, 128, 128, 3, 1, 1, False, "reflect"),
    (2, 6, 8, 8, 4, 2, 1, False, "reflect"),
    (4, 16, 128, 192, 3, 1, 1, True, "reflect"),    # tensor-core shapes (bf16 mode): 2 k-blocks, N=192
    (4, 16, 96, 96, 3, 2, 1, False, "reflect"),     # stride 2 (parity maps), Cin=96 (partial k-block)
    (2, 32, 64, 64, 3, 1, 1, False, "zero"),
    (8, 4, 256, 512, 3, 1, 1, False, "reflect"),    # 4x4 map, N tiles of 256
    (2, 16, 192, 24, 1, 1, 0, True, "reflect"),     # attention 1x1, N=24 (not a multiple of 16 -> SIMT)
    (2, 16, 192, 96, 1, 1, 0, True, "reflect"),
    (1, 128, 96, 16, 3, 1, 1, False, "reflect"),    # 128-pixel rows: halo kernel (taps of one dy share an activation box)
    (1, 128, 96, 96, 3, 1, 1, True, "zero"),        # halo kernel, zero padding through TMA OOB fill, 2 k-blocks
    (1, 256, 16, 32, 3, 1, 1, False, "reflect"),    # two 128-pixel tiles per image row
]


@pytest.mark.parametrize("dtype", DTYPES)
@pytest.mark.parametrize("case", CONV_CASES)
def test_conv(case, dtype):
    """ops.py:49-113: reflect/zero pad logic, SN weight, conv2d, bias; dgrad + wgrad."""
    from biggan_tensorflow_b200 import ops
    B, H, Cin, Cout, k, s, pad, bias, padtype = case
    po, ro = _opt(pad=padtype)
    x = np.random.default_rng(0).standard_normal((B, H, H, Cin)).astype(np.float32)
    run_case(lambda xs: ops.conv(xs["x"], Cout, po, kernel=k, stride=s, pad=pad, use_bias=bias),
             lambda rs, xs: R.conv(rs, xs["x"], Cout, ro, kernel=k, stride=s, pad=pad, use_bias=bias),
             {"x": x}, dtype)


@pytest.mark.parametrize("case", [(2, 16, 32, 3, 2, 1, True, "reflect"), (3, 9, 16, 3, 2, 1, False, "reflect"),
                                  (2, 8, 32, 3, 1, 1, True, "zero"), (2, 12, 96, 3, 1, 1, False, "reflect")])
def test_conv_rgb_im2col(case):
    """First D convs in tensor-core mode: the image is stored zero-padded to 16 channels (ops.rgb_pad()), conv() im2cols
    its 3 real channels (bg_im2col_rgb) and runs one 1x1 GEMM with K = 27 -> 32; dx comes back through bg_col2im_rgb."""
    from biggan_tensorflow_b200 import ops
    B, H, Cout, k, s, pad, bias, padtype = case
    po, ro = _opt(pad=padtype)
    x = np.random.default_rng(3).standard_normal((B, H, H, 3)).astype(np.float32)

    def prod(xs):
        xp = ops.pad_channels(xs["x"], 16)
        return ops.conv(xp, Cout, po, kernel=k, stride=s, pad=pad, use_bias=bias)

    run_case(prod, lambda rs, xs: R.conv(rs, xs["x"], Cout, ro, kernel=k, stride=s, pad=pad, use_bias=bias),
             {"x": x}, torch.bfloat16)


RGB_OUT_CASES = [(2, 8, 32, "reflect", True), (1, 128, 96, "reflect", True), (2, 33, 16, "reflect", False), (2, 30, 64, "zero", True),
                 (3, 2, 16, "reflect", False), (1, 61, 32, "zero", False)]


def rgb_out_body(case):
    """G_logit in tensor-core mode (BigGAN.py:570 + tanh :580): conv(relu-output x, 3 channels zero-padded to 16, 3x3, stride 1, pad 1)
    with the nine taps stacked along N -- bg_rgb_stack_weights, one 1x1 GEMM Ci -> 32 on the unpadded input, bg_rgb_tapsum_fwd
    (reflect / zero tap sum + tanh); backward through bg_rgb_tapsum_bwd, the 1x1 dgrad with the ReLU mask of x in its epilogue,
    the 1x1 wgrad and bg_rgb_unstack_wgrad.  Sizes cover one and several 30-pixel warp segments per row, ragged last segments,
    the 2-pixel minimum, 128-pixel rows."""
    from biggan_tensorflow_b200 import ops, _lib
    B, H, Cin, padtype, masked = case
    po, ro = _opt(pad=padtype)
    x = np.random.default_rng(5).standard_normal((B, H, H, Cin)).astype(np.float32)

    def prod(xs):
        xr = ops.relu(xs["x"])
        xr.relu_out = masked          # as after a fused BN + ReLU: the conv's dgrad applies the mask (relu's own backward repeats it)
        y = ops.conv(xr, 3, po, kernel=3, stride=1, pad=1, use_bias=False, scope="G_logit", _act=_lib.ACT_TANH, _pad_cout_to=16)
        assert y.shape[-1] == 16 and ops.c_real(y) == 3
        assert float(y.data[..., 3:].float().abs().max()) == 0.0
        return ops.slice_channels(y, 3)

    run_case(prod, lambda rs, xs: torch.tanh(R.conv(rs, R.relu(xs["x"]), 3, ro, kernel=3, stride=1, pad=1, use_bias=False, scope="G_logit")),
             {"x": x}, torch.bfloat16)


@pytest.mark.parametrize("case", RGB_OUT_CASES)
def test_conv_rgb_out_tap_stacked(case):
    from biggan_tensorflow_b200 import _lib
    n0 = _lib.launch_count
    rgb_out_body(case)
    assert _lib.launch_count > n0


DECONV_CASES = [(2, 4, 32, 16, 4, 2), (2, 8, 16, 16, 3, 1), (2, 4, 16, 8, 3, 2), (1, 4, 8, 8, 6, 2),   # deconv4/3/6
                (2, 5, 64, 64, 4, 2), (2, 4, 128, 64, 4, 2), (2, 4, 16, 16, 4, 1),
                (4, 8, 256, 128, 4, 2), (2, 16, 128, 128, 3, 1), (3, 4, 512, 512, 4, 2), (2, 32, 96, 96, 3, 1)]


@pytest.mark.parametrize("dtype", DTYPES)
@pytest.mark.parametrize("case", DECONV_CASES)
def test_deconv(case, dtype):
    """ops.py:116-139: conv2d_transpose SAME == adjoint of the SAME conv."""
    from biggan_tensorflow_b200 import ops
    B, H, Cin, Cout, k, s = case
    po, ro = _opt()
    x = np.random.default_rng(1).standard_normal((B, H, H, Cin)).astype(np.float32)
    run_case(lambda xs: ops.deconv(xs["x"], Cout, po, kernel=k, stride=s, use_bias=True),
             lambda rs, xs: R.deconv(rs, xs["x"], Cout, ro, kernel=k, stride=s, use_bias=True),
             {"x": x}, dtype)


@pytest.mark.parametrize("sn", [True, False])
@pytest.mark.parametrize("B,K,units", [(6, 52, 40), (6, 52, 200), (256, 148, 1536), (250, 300, 900), (3, 148, 96), (37, 1536, 1), (5, 100, 3)])
def test_fully_connected(B, K, units, sn):
    """ops.py:148-175 (SN path and tf.layers.dense path); the two large cases take the dedicated skinny-GEMM kernel
    (fc.cu): ragged row / column blocks, contraction longer than one shared-memory chunk."""
    from biggan_tensorflow_b200 import ops
    po, ro = _opt()
    po["conv"]["sn"] = sn
    ro["conv"]["sn"] = sn
    x = np.random.default_rng(2).standard_normal((B, 1, 1, K)).astype(np.float32)
    run_case(lambda xs: ops.fully_connected(xs["x"], units, po, scope="fc"),
             lambda rs, xs: R.fully_connected(rs, xs["x"], units, ro, scope="fc"),
             {"x": x}, torch.float32)


@pytest.mark.parametrize("dtype", DTYPES)
@pytest.mark.parametrize("shape", [(4, 4, 4, 64), (3, 8, 8, 32), (2, 16, 16, 96)])
def test_condition_batch_norm_relu(shape, dtype):
    """ops.py:611-643 followed by opt['act'] (ops.py:253-254): fused CBN + ReLU."""
    from biggan_tensorflow_b200 import ops, _lib
    po, ro = _opt()
    rng = np.random.default_rng(4)
    x = (rng.standard_normal(shape) * 1.5 + 0.3).astype(np.float32)
    z = rng.standard_normal((shape[0], 1, 1, 20)).astype(np.float32)
    run_case(lambda xs: ops.cond_bn(xs["x"], xs["z"], po, _act=_lib.ACT_RELU),
             lambda rs, xs: R.relu(R.cond_bn(rs, xs["x"], xs["z"], ro)),
             {"x": x, "z": z}, dtype, fp32_inputs=("z",), tol_scale=2.0)


@pytest.mark.parametrize("dtype", DTYPES)
def test_batch_norm(dtype):
    """ops.py:580-585 (tf.layers.batch_normalization training=True; moving stats updated)."""
    from biggan_tensorflow_b200 import ops
    po, ro = _opt()
    x = (np.random.default_rng(5).standard_normal((3, 8, 8, 32)) * 2 - 0.5).astype(np.float32)
    run_case(lambda xs: ops.bn(xs["x"], po), lambda rs, xs: R.bn(rs, xs["x"], ro), {"x": x}, dtype, tol_scale=2.0)


def test_batch_norm_inference_uses_population_stats():
    from biggan_tensorflow_b200 import ops
    po, ro = _opt(is_training=False)
    x = np.random.default_rng(6).standard_normal((2, 4, 4, 16)).astype(np.float32)
    z = np.random.default_rng(7).standard_normal((2, 1, 1, 12)).astype(np.float32)
    run_case(lambda xs: ops.cond_bn(xs["x"], xs["z"], po), lambda rs, xs: R.cond_bn(rs, xs["x"], xs["z"], ro),
             {"x": x, "z": z}, torch.float32)


def _sa_gamma(prod_store=None, ref_store=None, value=0.1):
    from biggan_tensorflow_b200.tf_shim import constant_initializer, get_variable, variable_scope
    if prod_store is not None:
        with variable_scope("self_attention"):
            get_variable("gamma", [1], initializer=constant_initializer(value))
    if ref_store is not None:
        with ref_store.variable_scope("self_attention"):
            ref_store.get_variable("gamma", [1], kind="const", value=value)


@pytest.mark.parametrize("dtype", DTYPES)
@pytest.mark.parametrize("shape,ch", [((2, 8, 8, 32), 32), ((2, 16, 16, 64), 64)])
def test_self_attention_2(shape, ch, dtype):
    """ops.py:467-492 with gamma = 0.1 so that the attention branch contributes."""
    from biggan_tensorflow_b200 import ops
    po, ro = _opt()
    x = np.random.default_rng(8).standard_normal(shape).astype(np.float32)

    def prod(xs):
        _sa_gamma(prod_store=True)
        return ops.self_attention_2(xs["x"], ch, po)

    def ref(rs, xs):
        _sa_gamma(ref_store=rs)
        return R.self_attention_2(rs, xs["x"], ch, ro)

    run_case(prod, ref, {"x": x}, dtype, tol_scale=2.0 if dtype == torch.float32 else 4.0)


def test_self_attention_gamma_zero_is_identity():
    """KAT: gamma is initialised to 0 (ops.py:486) => the block is the identity."""
    from biggan_tensorflow_b200 import ops
    po, _ = _opt()
    new_product_store(torch.float32)
    x = np.random.default_rng(9).standard_normal((2, 8, 8, 32)).astype(np.float32)
    y = ops.self_attention_2(dev_tensor(x), 32, po)
    assert np.array_equal(y.numpy(), x)


def test_self_attention_full():
    """ops.py:445-465 (API surface)."""
    from biggan_tensorflow_b200 import ops
    po, ro = _opt()
    x = np.random.default_rng(10).standard_normal((2, 4, 4, 32)).astype(np.float32)

    def prod(xs):
        _sa_gamma(prod_store=True)
        return ops.self_attention(xs["x"], 32, po)

    def ref(rs, xs):
        _sa_gamma(ref_store=rs)
        return R.self_attention(rs, xs["x"], 32, ro)

    run_case(prod, ref, {"x": x}, torch.float32)


@pytest.mark.parametrize("dtype", DTYPES)
def test_resblock_up_condition(dtype):
    """ops.py:250-266."""
    from biggan_tensorflow_b200 import ops
    po, ro = _opt()
    rng = np.random.default_rng(11)
    x = rng.standard_normal((2, 4, 4, 64)).astype(np.float32)
    z = rng.standard_normal((2, 1, 1, 24)).astype(np.float32)
    run_case(lambda xs: ops.resblock_up_condition(xs["x"], xs["z"], 32, po, use_bias=False, scope="resblock_up_4"),
             lambda rs, xs: R.resblock_up_condition(rs, xs["x"], xs["z"], 32, ro, use_bias=False, scope="resblock_up_4"),
             {"x": x, "z": z}, dtype, fp32_inputs=("z",), tol_scale=4.0 if dtype == torch.float32 else 7.5)   # bf16: see
    # tests/test_bf16_conditioning.py::test_bf16_projection_rounding_moves_the_small_resblock_case


@pytest.mark.parametrize("dtype", DTYPES)
@pytest.mark.parametrize("cin", [3, 32])
def test_resblock_down(cin, dtype):
    """ops.py:293-313."""
    from biggan_tensorflow_b200 import ops
    po, ro = _opt()
    x = np.random.default_rng(12).standard_normal((2, 8, 8, cin)).astype(np.float32)
    run_case(lambda xs: ops.resblock_down(xs["x"], 64, po, use_bias=False, scope="resblock_down_1"),
             lambda rs, xs: R.resblock_down(rs, xs["x"], 64, ro, use_bias=False, scope="resblock_down_1"),
             {"x": x}, dtype, tol_scale=3.0)


@pytest.mark.parametrize("dtype", DTYPES)
def test_resblock(dtype):
    """ops.py:187-198."""
    from biggan_tensorflow_b200 import ops
    po, ro = _opt()
    x = np.random.default_rng(13).standard_normal((2, 4, 4, 64)).astype(np.float32)
    run_case(lambda xs: ops.resblock(xs["x"], 64, po, use_bias=False),
             lambda rs, xs: R.resblock(rs, xs["x"], 64, ro, use_bias=False),
             {"x": x}, dtype, tol_scale=3.0)


def test_hinge_losses_and_flood():
    """ops.py:753-797, :799-840, :847-848 vs oracle + analytic KATs."""
    from biggan_tensorflow_b200 import ops
    from biggan_tensorflow_b200.tensor import backward
    rng = np.random.default_rng(14)
    for flood in (0.0, 0.1, 5.0):
        new_product_store(torch.float32)
        real = rng.standard_normal((16, 1)).astype(np.float32) * 2
        fake = rng.standard_normal((16, 1)).astype(np.float32) * 2
        tr, tf = dev_tensor(real), dev_tensor(fake)
        L = ops.discriminator_loss('hinge', real=tr, fake=tf, flood_level=flood)
        backward([(L, torch.ones(1, device="cuda"))])
        rr = torch.tensor(real, dtype=torch.float64, requires_grad=True)
        rf = torch.tensor(fake, dtype=torch.float64, requires_grad=True)
        Lr = R.discriminator_loss('hinge', rr, rf, flood)
        gr, gf = torch.autograd.grad(Lr, [rr, rf])
        assert_close(L.numpy(), Lr.detach().numpy(), 1e-5, "d_loss")
        assert_close(tr.grad.cpu().numpy(), gr.numpy(), 1e-5, "d real")
        assert_close(tf.grad.cpu().numpy(), gf.numpy(), 1e-5, "d fake")
        new_product_store(torch.float32)
        tf2 = dev_tensor(fake)
        G = ops.generator_loss('hinge', fake=tf2, real=None, flood_level=flood)
        backward([(G, torch.ones(1, device="cuda"))])
        rf2 = torch.tensor(fake, dtype=torch.float64, requires_grad=True)
        Gr = R.generator_loss('hinge', rf2, None, flood)
        (gg,) = torch.autograd.grad(Gr, [rf2])
        assert_close(G.numpy(), Gr.detach().numpy(), 1e-5, "g_loss")
        assert_close(tf2.grad.cpu().numpy(), gg.numpy(), 1e-5, "g fake")
    # KAT: real = fake = 0 -> D loss 2, G loss 0
    new_product_store(torch.float32)
    z = np.zeros((8, 1), np.float32)
    assert float(ops.discriminator_loss('hinge', dev_tensor(z), dev_tensor(z)).numpy()[0]) == 2.0
    assert float(ops.generator_loss('hinge', dev_tensor(z), None).numpy()[0]) == 0.0


def test_logistic_cls_loss():
    """utils.py:366-369."""
    from biggan_tensorflow_b200 import ops
    from biggan_tensorflow_b200.tensor import backward
    rng = np.random.default_rng(15)
    new_product_store(torch.float32)
    logits = (rng.standard_normal((6, 10)) * 3).astype(np.float32)
    labels = np.eye(10, dtype=np.float32)[rng.integers(0, 10, 6)]
    w = rng.random(10).astype(np.float32) + 0.5
    tl = dev_tensor(logits)
    fn = ops.cls_loss_fn('logistic', torch.from_numpy(w).cuda())
    L = 5.0 * fn(torch.from_numpy(labels).cuda(), tl)
    backward([(L, torch.ones(1, device="cuda"))])
    rl = torch.tensor(logits, dtype=torch.float64, requires_grad=True)
    Lr = 5.0 * R.cls_loss_logistic(torch.tensor(labels, dtype=torch.float64), rl, torch.tensor(w, dtype=torch.float64))
    (g,) = torch.autograd.grad(Lr, [rl])
    assert_close(L.numpy(), Lr.detach().numpy(), 1e-5, "cls loss")
    assert_close(tl.grad.cpu().numpy(), g.numpy(), 1e-5, "cls grad")


def test_adam_ema_matches_tf_formula():
    """tf.train.AdamOptimizer + MovingAverageOptimizer (BigGAN.py:923-926)."""
    from biggan_tensorflow_b200._lib import call
    from biggan_tensorflow_b200.tensor import stream
    rng = np.random.default_rng(16)
    n = 1003
    p = rng.standard_normal(n).astype(np.float32)
    m = np.zeros(n, np.float32)
    v = np.zeros(n, np.float32)
    ema = p.copy()
    tp, tm, tv, te = [torch.from_numpy(a.copy()).cuda() for a in (p, m, v, ema)]
    b1, b2, lr, d = 0.0, 0.9, 5e-5, 0.999
    p64, m64, v64, e64 = [a.astype(np.float64) for a in (p, m, v, ema)]
    for t in range(1, 4):
        g = rng.standard_normal(n).astype(np.float32)
        lr_t = lr * np.sqrt(1 - b2 ** t) / (1 - b1 ** t)
        call("bg_adam_ema_step", tp, torch.from_numpy(g).cuda(), tm, tv, te, n, float(lr_t), None, b1, b2, 1e-8, d, 1.0, stream())
        m64 = b1 * m64 + (1 - b1) * g
        v64 = b2 * v64 + (1 - b2) * g.astype(np.float64) ** 2
        p64 = p64 - lr_t * m64 / (np.sqrt(v64) + 1e-8)
        e64 = d * e64 + (1 - d) * p64
    torch.cuda.synchronize()
    assert_close(tp.cpu().numpy(), p64, 1e-6, "param")
    assert_close(te.cpu().numpy(), e64, 1e-6, "ema")
    assert_close(tv.cpu().numpy(), v64, 1e-6, "v")


def _tc_counters():
    import ctypes
    from biggan_tensorflow_b200 import _lib
    a, b = ctypes.c_int(0), ctypes.c_int(0)
    _lib.lib.bg_debug_counters(ctypes.byref(a), ctypes.byref(b))
    return a.value, b.value


def test_tensor_core_path_is_taken_in_bf16_mode():
    """The bf16 conv path must run on the tcgen05 kernels (fwd, dgrad, wgrad), not on the SIMT fallback,
    and must agree with the SIMT kernels on the same bf16 inputs."""
    from biggan_tensorflow_b200 import _lib, ops
    assert _lib.lib.bg_has_tensor_core_path() == 1
    po, _ = _opt()
    x = np.random.default_rng(20).standard_normal((4, 16, 16, 128)).astype(np.float32)
    dy = np.random.default_rng(21).standard_normal((4, 16, 16, 192)).astype(np.float32)
    res = {}
    for simt in (0, 1):
        new_product_store(torch.bfloat16)
        ops.set_force_simt(simt)
        try:
            f0, w0 = _tc_counters()
            tx = dev_tensor(x, torch.bfloat16)
            y = ops.conv(tx, 192, po, kernel=3, stride=1, pad=1, use_bias=False)
            seed_backward(y, dy)
            f1, w1 = _tc_counters()
        finally:
            ops.set_force_simt(0)
        from biggan_tensorflow_b200.tf_shim import get_store
        res[simt] = (y.numpy(), tx.grad.float().cpu().numpy(), get_store().vars["conv_0/kernel"].grad.cpu().numpy())
        if simt == 0:
            assert f1 - f0 == 2 and w1 - w0 == 1, (f0, f1, w0, w1)      # fwd + dgrad, wgrad
        else:
            assert f1 == f0 and w1 == w0
    for a, b, what in zip(res[0], res[1], ("y", "dx", "dw")):
        assert_close(a, b, 1e-2, "tcgen05 vs SIMT " + what)
