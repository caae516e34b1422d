"""The Python glue of the SURVEY 8(f) operators (`ops.lrelu / prelu / avg_pooling / up_sample / depth_to_space`,
`utils.orthogonal_regularizer(_fc)`) executed on CPU tensors: for this test only, the package's device / stream helpers
are pointed at the CPU and `_lib.lib` at a host build of `csrc/next_ops.cu` (tests/host_emul/host_lib.cpp: the real
kernel source under the emulation prelude + loop stand-ins for bg_add / bg_axpby / bg_bgemm).  It checks what a GPU
run would check about the glue -- argument order into the C ABI, variable creation and naming, the tape and gradient
accumulation -- against the reference-generated vectors.  The product package itself has no CPU path: nothing outside
this test can select the host library."""
import os
import sys

import numpy as np
import pytest
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "golden"))
import ref_cases as C                       # noqa: E402
sys.path.insert(0, os.path.join(HERE, "host_emul"))
import host_env                             # noqa: E402

ROOT = os.path.dirname(HERE)
STANDALONE = {"prelu", "lrelu", "upconv_nn", "downconv_pool_only"}           # cases that need no conv kernel


@pytest.fixture()
def host_product(host_so, monkeypatch):
    from biggan_tensorflow_b200 import ops, tensor, tf_shim
    host_env.activate(host_so, monkeypatch.setattr)
    st = tf_shim.set_store(tf_shim.VariableStore())
    ops.set_compute_dtype(torch.float32)
    tensor.clear_tape()
    yield st
    tf_shim.set_store(tf_shim.VariableStore())


def _tensor(a, requires_grad=True):
    from biggan_tensorflow_b200.tensor import Tensor
    return Tensor(torch.from_numpy(np.ascontiguousarray(a, dtype=np.float32)), requires_grad=requires_grad)


@pytest.mark.parametrize("case", [c for c in C.NEXT_OP_CASES if c["name"] in STANDALONE], ids=lambda c: c["name"])
def test_glue_of_standalone_ops_matches_reference_source(host_product, case):
    from biggan_tensorflow_b200 import ops
    from biggan_tensorflow_b200.tensor import backward
    g = np.load(os.path.join(HERE, "golden", "ref_ops_next.npz"))
    p = case["name"]
    st = host_product
    if case["var_init"]:
        st.load({k: np.asarray(v, np.float32) for k, v in case["var_init"].items()})
    acts = {"relu": ops.relu, "prelu": ops.prelu, "lrelu": lambda x: ops.lrelu(x, 0.2)}
    opt = C.make_opt(acts, case["opt"])
    xs = {k: _tensor(v) for k, v in C.case_inputs(case).items()}
    if case["fn"] in ("prelu", "lrelu"):
        y = getattr(ops, case["fn"])(*xs.values())
    else:
        y = getattr(ops, case["fn"])(*xs.values(), **case["kw"], opt=opt)
    backward([(y, torch.from_numpy(C.case_dy(case, tuple(y.shape))))])
    C.check(g, p + "/out", y.numpy(), 1e-6)
    C.check(g, p + "/gin/x", xs["x"].grad.numpy(), 1e-6)
    assert sorted(st.vars) == sorted(k.split("/varsum/")[1] for k in g.files if k.startswith(p + "/varsum/"))
    for key in (k for k in g.files if k.startswith(p + "/gvar/")):
        C.check(g, key, st.vars[key.split("/gvar/")[1]].grad.numpy(), 1e-6)


def test_glue_gradient_accumulation_and_padded_prelu(host_product):
    """x feeding two consumers (grad_target's second-contribution path) and prelu on a zero-padded image tensor."""
    from biggan_tensorflow_b200 import ops
    from biggan_tensorflow_b200.tensor import backward
    rng = np.random.default_rng(0)
    x = rng.standard_normal((2, 4, 4, 6)).astype(np.float32)
    t = _tensor(x)
    y = ops.add(ops.avg_pooling(ops.up_sample(t)), ops.lrelu(t, 0.2))        # avg_pool(up_sample(x)) == x
    backward([(y, torch.ones(2, 4, 4, 6))])
    assert np.allclose(y.numpy(), x + np.maximum(x, 0.2 * x), atol=1e-6)
    assert np.allclose(t.grad.numpy(), 1.0 + np.where(x > 0, 1.0, 0.2), atol=1e-6)
    d = ops.depth_to_space(_tensor(rng.standard_normal((1, 2, 2, 8)).astype(np.float32)), 2)
    assert d.shape == (1, 4, 4, 2)
    img = rng.standard_normal((2, 3, 3, 3)).astype(np.float32)
    host_product.load({"prelu/alpha": np.array([0.5, -1.0, 0.25], np.float32)})
    padded = np.concatenate([img, np.zeros((2, 3, 3, 13), np.float32)], axis=-1)
    tp = _tensor(padded)
    tp.aux = {"c_real": 3}
    yp = ops.prelu(tp)
    backward([(yp, torch.ones(2, 3, 3, 16))])
    a = np.array([0.5, -1.0, 0.25], np.float32)
    assert np.allclose(yp.numpy()[..., :3], np.where(img > 0, img, a * img), atol=1e-6) and not yp.numpy()[..., 3:].any()
    assert np.allclose(host_product.vars["prelu/alpha"].grad.numpy(), np.minimum(img, 0).sum(axis=(0, 1, 2)), atol=1e-5)


@pytest.mark.parametrize("kind", ["ortho", "ortho_cosine"])
def test_glue_of_the_orthogonal_regularizers_matches_reference_source(host_product, kind):
    from biggan_tensorflow_b200 import utils
    from biggan_tensorflow_b200.tensor import backward
    g = np.load(os.path.join(HERE, "golden", "ref_ops.npz"))
    for wkey, lkey, gkey, fn in (("w4", "loss4", "g4", utils.orthogonal_regularizer), ("w2", "loss2", "g2", utils.orthogonal_regularizer_fc)):
        w = _tensor(g[kind + "/" + wkey])
        L = fn(1e-4, type=kind)(w)
        backward([(L, torch.full((1,), 3.0))])                                # upstream gradient 3: scaled on the device
        C.check(g, kind + "/" + lkey, L.numpy().reshape(()), 2e-5)
        C.check(g, kind + "/" + gkey, w.grad.numpy() / 3.0, 2e-5)
        # the Trainer's path: loss accumulated into a buffer, gradient added to an existing arena slice
        loss = torch.zeros(1)
        dW = torch.zeros_like(w.data)
        for _ in range(2):                                                    # += both times
            fn(1e-4, type=kind).evaluate(w.data, loss, dW=dW, grad_scale=0.5)
        C.check(g, kind + "/" + lkey, loss.numpy().reshape(()) / 2.0, 2e-5)
        C.check(g, kind + "/" + gkey, dW.numpy(), 2e-5)
    with pytest.raises(ValueError, match="Unknown regularization method."):
        utils.orthogonal_regularizer(1e-4, type="l3")(_tensor(np.ones((2, 2), np.float32)))


def test_glue_of_the_renorm_affine_map(host_product):
    """ops._renorm_affine: gamma' = r gamma, beta' = beta + d gamma with r, d constant -> dgamma = r dgamma' + d dbeta', dbeta = dbeta'."""
    from biggan_tensorflow_b200 import ops
    from biggan_tensorflow_b200.tensor import backward
    rng = np.random.default_rng(3)
    gamma, beta = _tensor(rng.standard_normal((4, 9))), _tensor(rng.standard_normal((4, 9)))
    r, d = torch.from_numpy(rng.random(9).astype(np.float32) + 0.5), torch.from_numpy(rng.standard_normal(9).astype(np.float32))
    g2, b2 = ops._renorm_affine(gamma, beta, r, d)
    dg2, db2 = torch.from_numpy(rng.standard_normal((4, 9)).astype(np.float32)), torch.from_numpy(rng.standard_normal((4, 9)).astype(np.float32))
    backward([(g2, dg2), (b2, db2)])
    assert np.allclose(g2.numpy(), r.numpy() * gamma.numpy(), atol=1e-6) and np.allclose(b2.numpy(), beta.numpy() + d.numpy() * gamma.numpy(), atol=1e-6)
    assert np.allclose(gamma.grad.numpy(), r.numpy() * dg2.numpy() + d.numpy() * db2.numpy(), atol=1e-5)
    assert np.allclose(beta.grad.numpy(), db2.numpy(), atol=1e-6)


@pytest.mark.parametrize("name", ["cbrn_fresh", "cbrn_warm", "cbrn_shared", "cbrn_inference"])
def test_glue_of_condition_batch_renorm_matches_reference_source(host_product, monkeypatch, name):
    """ops.condition_batch_renorm (ops.py:645-715) as shipped -- statistics pass, bg_renorm_factors, the affine map, then the
    conditional batch-norm kernel set (loop stand-ins here) -- against the reference-generated vectors: output, d/dx, and all
    state (pop_mean / pop_var, renorm_mean / renorm_var, renorm_weight).  The two SN projections are oracle stand-ins."""
    import torch as _torch
    from biggan_tensorflow_b200 import ops
    from biggan_tensorflow_b200.tensor import Tensor, backward
    from oracle import ref_ops as R
    case = [c for c in C.NEXT_OP_CASES if c["name"] == name][0]
    g = np.load(os.path.join(HERE, "golden", "ref_ops_next.npz"))
    st = host_product
    st.load({k: np.asarray(v, np.float32) for k, v in case["var_init"].items()})
    for k in list(st.vars):                                            # state variables (load() guesses by name)
        st.vars[k].trainable = st.vars[k].requires_grad = False
    rs = R.RefStore(_torch.float64)

    def fc(z, units, opt, use_bias=True, lrmul=1.0, scope='fully_0'):
        from biggan_tensorflow_b200 import tf_shim
        rs.scope = list(tf_shim.get_store().scope)
        y = R.fully_connected(rs, z.data.double(), units, opt, use_bias, lrmul, scope)
        return Tensor(y.detach().float().contiguous(), requires_grad=True)

    monkeypatch.setattr(ops, "fully_connected", fc)
    opt = C.make_opt({"relu": ops.relu}, case["opt"])
    xs = {k: _tensor(v) for k, v in C.case_inputs(case).items()}
    y = ops.cond_bn(xs["x"], xs["z"], opt=opt)
    backward([(y, _torch.from_numpy(C.case_dy(case, tuple(y.shape))))])
    C.check(g, name + "/out", y.numpy(), 2e-5)
    C.check(g, name + "/gin/x", xs["x"].grad.numpy(), 2e-5)
    for key in (k for k in g.files if k.startswith(name + "/state/")):
        n = key.split("/state/")[1]
        if n.endswith("/u"):
            continue                                                   # the projections' spectral-norm state lives in the stand-in
        C.check(g, key, st.vars[n].data.numpy(), 2e-5)
