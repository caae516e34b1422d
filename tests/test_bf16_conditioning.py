"""Why two gradients of the attention block get a wider bf16 tolerance: rounding the activations of the
reference computation itself to bf16 (in float64 arithmetic, no CUDA involved) already moves d/d f_conv/kernel
by ~1e-1 relative L2, because bf16 rounding flips the 2x2 max-pool argmax of the key branch.  The CUDA tests
therefore allow 1.5e-1 on the max-pooled branches (f_conv, h_conv kernels) and keep 2e-2 elsewhere."""
import numpy as np
import torch

from oracle import ref_ops as R


class _RoundBF16(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x):
        return x.float().bfloat16().double()

    @staticmethod
    def backward(ctx, g):
        return g.float().bfloat16().double()


def _grads(round_):
    rb = _RoundBF16.apply if round_ else (lambda t: t)
    rs = R.RefStore(torch.float64)
    with rs.variable_scope("self_attention"):
        rs.get_variable("gamma", [1], kind="const", value=0.1)
    x = torch.tensor(np.random.default_rng(8).standard_normal((2, 16, 16, 64)), dtype=torch.float64)
    opt = {"is_training": True, "self_attention_bias": True, "conv": {"padding_type": "reflect", "sn": True}}
    ch = 64
    xx = rb(x)
    with rs.variable_scope("self_attention"):
        f = R.max_pooling(rb(R.conv(rs, xx, ch // 8, kernel=1, stride=1, opt=opt, scope="f_conv", use_bias=True)))
        g = rb(R.conv(rs, xx, ch // 8, kernel=1, stride=1, opt=opt, scope="g_conv", use_bias=True))
        h = R.max_pooling(rb(R.conv(rs, xx, ch // 2, kernel=1, stride=1, opt=opt, scope="h_conv", use_bias=True)))
        beta = torch.softmax(R.hw_flatten(g) @ R.hw_flatten(f).transpose(1, 2), -1)
        o = rb(beta @ R.hw_flatten(h)).reshape(2, 16, 16, ch // 2)
        o = rb(R.conv(rs, o, ch, kernel=1, stride=1, opt=opt, scope="attn_conv", use_bias=True))
        y = rs.vars["self_attention/gamma"] * o + xx
    dy = torch.tensor(np.random.default_rng(1000).standard_normal(y.shape))
    names = [n for n in rs.vars if rs.trainable.get(n)]
    return dict(zip(names, torch.autograd.grad((y * dy).sum(), [rs.vars[n] for n in names])))


def test_bf16_rounding_alone_moves_the_key_branch_gradient():
    a, b = _grads(False), _grads(True)
    rel = {n: float((a[n] - b[n]).norm() / a[n].norm().clamp_min(1e-30)) for n in a if not n.endswith("f_conv/bias")}
    assert rel["self_attention/f_conv/kernel"] > 4e-2          # inherent, not a kernel error
    assert rel["self_attention/f_conv/kernel"] < 1.5e-1
    assert rel["self_attention/g_conv/kernel"] < 2e-2 and rel["self_attention/attn_conv/kernel"] < 2e-2


def test_bf16_storage_alone_moves_whole_network_generator_gradients():
    """The G update back-propagates through all of D and all of G (~60 bf16-stored tensors).  Rounding just
    the stored activations/gradients of the REFERENCE computation to bf16 (float64 arithmetic, oracle hook
    `ACT_ROUND`) already moves the earliest generator gradients by ~1e-1 relative L2, while the loss moves by
    ~1e-3: the bf16-mode whole-network gradient tolerance (tests/test_step_gpu.py) is set from this level."""
    from oracle import ref_model as M
    kw = dict(img_size=32, ch=32, batch_size=8, n_labels=10, sa_gamma_init=0.1)
    cfg = M.Config(**kw)
    b = M.synthetic_batch(cfg, B=8)
    out = {}
    for rounded in (False, True):
        R.ACT_ROUND = _RoundBF16.apply if rounded else None
        try:
            st = R.RefStore(torch.float64)
            M.d_update(st, cfg, b["real"], b["labels"], b["z_d"], b["labels"], opt_state=M.AdamState(cfg.d_lr, 0.0, 0.9))
            loss, grads, _ = M.g_update(st, cfg, b["z_g"], b["labels"], apply=False)
        finally:
            R.ACT_ROUND = None
        out[rounded] = (loss, grads)
    rel = {n: float((out[True][1][n] - g).norm() / g.norm().clamp_min(1e-30)) for n, g in out[False][1].items()}
    assert abs(out[True][0] - out[False][0]) <= 5e-3 * abs(out[False][0])
    assert 2e-2 < rel["generator/dense1/kernel"] < 3e-1
    assert max(v for n, v in rel.items() if not n.endswith("f_conv/bias")) < 4e-1
    print({k: round(v, 3) for k, v in rel.items() if "dense" in k or "G_logit" in k})


def test_bf16_projection_rounding_moves_the_small_resblock_case():
    """In bf16 mode the conditional-BatchNorm gamma/beta projections (fully_connected, ops.py:148-175) run on the tensor
    cores with bf16 operands like every other dense contraction.  Rounding ONLY the projection's operands and result
    in the FORWARD pass of the REFERENCE computation (float64 arithmetic) moves the block output by 2e-3 but the
    gradients of the small resblock test case (batch 2, 4x4 -> 8x8, 32 BatchNorm samples per channel) by 5e-2..1e-1
    relative L2: a gamma/beta perturbation shifts a whole (sample, channel) slice coherently, so the ReLU masks behind it
    flip together instead of averaging out.  Rounding in the backward pass alone moves them by < 5e-3.
    tests/test_ops_gpu.py::test_resblock_up_condition therefore allows 1.5e-1 in bf16 mode (fp32 mode keeps 4e-4)."""
    class _FwdOnly(torch.autograd.Function):
        @staticmethod
        def forward(ctx, x):
            return x.float().bfloat16().double()

        @staticmethod
        def backward(ctx, g):
            return g

    class _BwdOnly(torch.autograd.Function):
        @staticmethod
        def forward(ctx, x):
            return x.clone()

        @staticmethod
        def backward(ctx, g):
            return g.float().bfloat16().double()

    def run(rb):
        orig = R.fully_connected
        if rb is not None:
            def fc(store, x, units, opt, use_bias=True, lrmul=1.0, scope="fully_0"):
                with store.variable_scope(scope):
                    x = x.reshape(x.shape[0], -1)
                    w = store.get_variable("kernel", [x.shape[-1], units])
                    b = store.get_variable("bias", [units], kind="const", value=0.0)
                    return rb(rb(x) @ rb(R.spectral_norm(store, w)) + b)
            R.fully_connected = fc
        try:
            rs = R.RefStore(torch.float64)
            ro = {"is_training": True, "bn_in_d": False, "act": R.relu, "upsampling_method": "deconv4", "g_conv": "deconv3",
                  "downsampling_method": "strided_conv3", "self_attention_bias": True, "fc_regularizer": None,
                  "bn": {"type": "batch_norm", "momentum": 0.98}, "conv": {"padding_type": "reflect", "sn": True}}
            rng = np.random.default_rng(11)
            x = torch.tensor(rng.standard_normal((2, 4, 4, 64)).astype(np.float32)).double()
            z = torch.tensor(rng.standard_normal((2, 1, 1, 24)).astype(np.float32)).double()
            y = R.resblock_up_condition(rs, x, z, 32, ro, use_bias=False, scope="resblock_up_4")
            dy = torch.tensor(np.random.default_rng(1000).standard_normal(y.shape))
            names = [n for n in rs.vars if rs.trainable.get(n)]
            return y.detach(), dict(zip(names, torch.autograd.grad((y * dy).sum(), [rs.vars[n] for n in names])))
        finally:
            R.fully_connected = orig

    y0, a = run(None)
    y1, f = run(_FwdOnly.apply)
    _, b = run(_BwdOnly.apply)
    rel = lambda u, n: float((a[n] - u[n]).norm() / a[n].norm().clamp_min(1e-30))
    assert float((y1 - y0).norm() / y0.norm()) < 5e-3
    proj = [n for n in a if "batch_norm" in n and n.endswith("kernel")]
    assert len(proj) == 4
    assert max(rel(f, n) for n in proj) > 4e-2                  # inherent to the case, not a kernel error
    assert max(rel(f, n) for n in a) < 1.5e-1
    assert max(rel(b, n) for n in a) < 5e-3
