"""Known-answer tests that pin the CPU oracle (the reference ships no tests / golden vectors and TensorFlow 1.x
cannot run here: PARITY UNPINNED -- these analytic facts, SURVEY 8c (3)-(4), are what pins it instead)."""
import numpy as np
import torch

from oracle import ref_model as M
from oracle import ref_ops as R

OPT = {"is_training": True, "bn_in_d": False, "act": R.relu, "upsampling_method": "deconv4", "g_conv": "deconv3",
       "downsampling_method": "strided_conv3", "self_attention_bias": True, "bn": {"type": "batch_norm", "momentum": 0.98},
       "conv": {"padding_type": "reflect", "sn": True}, "fc_regularizer": None}


def test_hinge_known_answers():
    z = torch.zeros(8, 1, dtype=torch.float64)
    assert float(R.discriminator_loss("hinge", z, z)) == 2.0          # ops.py:789-790
    assert float(R.generator_loss("hinge", z, None)) == 0.0           # ops.py:833
    L = R.discriminator_loss("hinge", z + 3.0, z - 3.0, flood_level=0.1)
    assert abs(float(L) - (abs(0.0 - 0.1) + 0.1)) < 1e-12             # flood |L-f|+f, ops.py:847-848


def test_spectral_norm_rank1_sigma_is_exact_and_norm_is_one():
    rng = np.random.default_rng(0)
    a, b = rng.standard_normal(30), rng.standard_normal(12)
    st = R.RefStore(torch.float64)
    w = torch.tensor(np.outer(a, b))
    wn = R.spectral_norm(st, w)
    assert abs(st.last_sigma - np.linalg.norm(a) * np.linalg.norm(b)) < 1e-10
    assert abs(np.linalg.svd(wn.numpy(), compute_uv=False)[0] - 1.0) < 1e-10


def test_spectral_norm_sigma_equals_norm_of_vW_and_closed_form_gradient():
    """sigma == |v_hat W| and dW = G/sigma - (<G,W>/sigma^2) v_hat^T u_hat."""
    rng = np.random.default_rng(1)
    W = torch.tensor(rng.standard_normal((72, 16)), requires_grad=True)
    st = R.RefStore(torch.float64)
    u0 = st.get_variable("u", [1, 16], kind="normal", trainable=False).clone()
    wn = R.spectral_norm(st, W)
    G = torch.tensor(rng.standard_normal((72, 16)))
    (g,) = torch.autograd.grad((wn * G).sum(), [W])
    v = R.l2_normalize(u0 @ W.detach().t())
    uu = R.l2_normalize(v @ W.detach())
    sigma = float((v @ W.detach() @ uu.t()))
    assert abs(sigma - float((v @ W.detach()).norm())) < 1e-12
    closed = G / sigma - ((G * W.detach()).sum() / sigma ** 2) * (v.t() @ uu)
    assert float((g - closed).abs().max()) < 1e-12
    assert torch.allclose(st.vars["u"], uu)                            # u.assign(u_hat), ops.py:743


def test_spectral_norm_evaluated_once_per_phase():
    st = R.RefStore(torch.float64)
    w = torch.tensor(np.random.default_rng(2).standard_normal((20, 8)))
    a = R.spectral_norm(st, w)
    u1 = st.vars["u"].clone()
    b = R.spectral_norm(st, w)                                         # same phase: cached, u unchanged
    assert a is b and torch.equal(u1, st.vars["u"])
    st.begin_phase()
    R.spectral_norm(st, w)
    assert not torch.equal(u1, st.vars["u"])                           # next phase advances u


def test_deconv_is_the_adjoint_of_the_same_conv():
    """tf.nn.conv2d_transpose(SAME) == input-gradient of the SAME conv, k in {3,4,6}, s in {1,2}."""
    rng = np.random.default_rng(3)
    for k, s in ((4, 2), (3, 1), (3, 2), (6, 2), (4, 1)):
        st = R.RefStore(torch.float64)
        x = torch.tensor(rng.standard_normal((2, 5, 5, 6)))
        y = R.deconv(st, x, 4, OPT, kernel=k, stride=s, use_bias=False)
        wn = st.sn_cache["deconv_0/u"]                                  # [k,k,Cout,Cin]
        yin = torch.tensor(rng.standard_normal((2, 5 * s, 5 * s, 4)), requires_grad=True)
        pt, pb = R.tf_same_pad(5 * s, k, s)
        xp = torch.nn.functional.pad(yin.permute(0, 3, 1, 2), (pt, pb, pt, pb))
        fwd = torch.nn.functional.conv2d(xp, wn.permute(3, 2, 0, 1), stride=s)          # SAME conv, HWIO = [k,k,Cout->in,Cin->out]
        (adj,) = torch.autograd.grad((fwd * x.permute(0, 3, 1, 2)).sum(), [yin])
        assert float((adj - y).abs().max()) < 1e-12, (k, s)


def test_reflect_pad_conv_with_delta_kernel_is_identity():
    st = R.RefStore(torch.float64)
    x = torch.tensor(np.random.default_rng(4).standard_normal((1, 6, 6, 3)))
    w = np.zeros((3, 3, 3, 3))
    for c in range(3):
        w[1, 1, c, c] = 1.0
    st.vars["conv_0/kernel"] = torch.tensor(w)
    y = R.conv(st, x, 3, OPT, kernel=3, stride=1, pad=1, use_bias=False)
    assert float((y * st.last_sigma - x).abs().max()) < 1e-12


def test_condition_batch_norm_statistics():
    """CBN output: per-channel batch mean of (y - beta_b)/gamma_b is 0 with unit variance (ops.py:630-638)."""
    st = R.RefStore(torch.float64)
    rng = np.random.default_rng(5)
    x = torch.tensor(rng.standard_normal((6, 4, 4, 8)) * 3 + 1)
    z = torch.tensor(rng.standard_normal((6, 1, 1, 5)))
    y = R.condition_batch_norm(st, x, z, OPT)
    st.begin_phase()
    beta = R.fully_connected(st, z, 8, OPT, scope="batch_norm/beta").reshape(6, 1, 1, 8)
    # (second evaluation advanced u; recompute the normalised activations directly instead)
    xh = (x - x.mean(dim=(0, 1, 2))) * torch.rsqrt(((x - x.mean(dim=(0, 1, 2))) ** 2).mean(dim=(0, 1, 2)) + 1e-5)
    assert float(xh.mean(dim=(0, 1, 2)).abs().max()) < 1e-12
    assert float((xh.var(dim=(0, 1, 2), unbiased=False) - 1).abs().max()) < 1e-4
    assert abs(float(st.vars["batch_norm/pop_mean"][0]) - 0.02 * float(x.mean(dim=(0, 1, 2))[0])) < 1e-12   # EMA, ops.py:631
    assert y.shape == x.shape and beta.shape == (6, 1, 1, 8)


def test_attention_softmax_rows_and_identity_at_gamma_zero():
    st = R.RefStore(torch.float64)
    x = torch.tensor(np.random.default_rng(6).standard_normal((2, 4, 4, 16)))
    y = R.self_attention_2(st, x, 16, OPT)
    assert torch.equal(y, x)                                           # gamma initialised to 0 (ops.py:486)


def test_finite_difference_gradient_through_discriminator_and_generator():
    """float64 finite differences vs autograd.  Differentiated w.r.t. network INPUTS: the
    spectral-norm weights then stay fixed, so stop_gradient(u_hat, v_hat) (ops.py:738-739) does not enter."""
    cfg = M.Config(img_size=32, ch=8, batch_size=2, n_labels=4, sa_gamma_init=0.1)
    b = M.synthetic_batch(cfg, B=2)

    def d_of(img):
        s = R.RefStore(torch.float64)                      # fresh store: same deterministic weights and u
        return R.discriminator_loss("hinge", M.discriminator(s, cfg, img)["real"], M.discriminator(s, cfg, img * 0.5)["real"])

    x = torch.tensor(b["real"], dtype=torch.float64, requires_grad=True)
    (g,) = torch.autograd.grad(d_of(x), [x])
    for idx in ((0, 3, 4, 1), (1, 17, 30, 2)):
        e = torch.zeros_like(x)
        e[idx] = 1e-6
        fd = float(d_of(x.detach() + e) - d_of(x.detach() - e)) / 2e-6
        assert abs(fd - float(g[idx])) <= 1e-6 * max(1.0, abs(fd)), (idx, fd, float(g[idx]))

    def g_of(z):
        s = R.RefStore(torch.float64)
        return (M.generator(s, cfg, z, torch.tensor(b["labels"], dtype=torch.float64)) ** 2).sum()

    z = torch.tensor(b["z_g"], dtype=torch.float64, requires_grad=True)
    (gz,) = torch.autograd.grad(g_of(z), [z])
    for idx in ((0, 0, 0, 5), (1, 0, 0, 200)):
        e = torch.zeros_like(z)
        e[idx] = 1e-6
        fd = float(g_of(z.detach() + e) - g_of(z.detach() - e)) / 2e-6
        assert abs(fd - float(gz[idx])) <= 1e-5 * max(1.0, abs(fd)), (idx, fd, float(gz[idx]))


def test_init_matches_between_oracle_and_product_spec():
    """Both sides initialise a variable from its name alone: identical arrays without any transfer."""
    from biggan_tensorflow_b200 import tf_shim
    for name, shape in (("generator/dense1/kernel", [40, 24]), ("discriminator/D_logit/u", [1, 1])):
        kind = "normal" if name.endswith("/u") else "trunc_normal"
        a = R.init_array(name, shape, kind)
        init = tf_shim.random_normal_initializer() if kind == "normal" else tf_shim.truncated_normal_initializer(0.0, 0.02)
        assert np.array_equal(a, init(name, shape))
    assert np.abs(R.init_array("x/kernel", [1000], "trunc_normal")).max() <= 0.04 + 1e-9      # truncated at 2 sigma
