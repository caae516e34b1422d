"""Wiring check of `model.py` and of the composite operators of `ops.py` (resblock_*, upconv / g_conv / downconv with
their variants, the activation / bn_in_d branches) WITHOUT a GPU: for this test only, the LEAF operators (conv, deconv,
fully_connected, the batch norms, attention, activations, pooling, adds) are replaced by float64 stand-ins built on the
oracle, while `model.generator` / `model.discriminator` and the composites run as shipped.  The forward results and the
set of variables they create are compared with the vectors produced by the reference's own generator() /
discriminator() (tests/golden/ref_net*.npz) -- for the canonical 64 / 128 / 256 / 512 pixel networks and for the SURVEY 8(f)
option sets (CLI defaults: prelu + regulariser + flood + DiffAugment; variants: lrelu, resize_conv, conv3,
resize_conv1, bn_in_d).  What this does not cover: the kernels themselves and the backward tape (GPU tests)."""
import os
import sys

import numpy as np
import pytest
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "golden"))
import ref_cases as C                       # noqa: E402
from oracle import ref_model as M           # noqa: E402
from oracle import ref_ops as R             # noqa: E402


@pytest.fixture()
def leaf_stand_ins(monkeypatch):
    from biggan_tensorflow_b200 import _lib, model, ops, tensor, tf_shim
    from biggan_tensorflow_b200.tensor import Tensor
    rs = R.RefStore(torch.float64, params={"generator/self_attention/gamma": np.array([0.1], np.float32),
                                           "discriminator/self_attention/gamma": np.array([0.1], np.float32)})
    rs.begin_phase()
    st = tf_shim.set_store(tf_shim.VariableStore())
    for mod in (tensor, tf_shim):
        monkeypatch.setattr(mod, "device", lambda: torch.device("cpu"))
    ops.set_compute_dtype(torch.float32)

    def sync():
        """Mirror the product store's scope stack into the oracle store (and, like a scope exit would, forget the
        default-name counters of every scope that has been left)."""
        rs.scope = list(tf_shim.get_store().scope)
        cur = rs.full_name()
        for k in [k for k in rs.default_names if k and cur != k and not cur.startswith(k + "/")]:
            del rs.default_names[k]

    def T(t, like=None):
        y = Tensor(t)
        if like is not None and isinstance(like.aux, dict) and "c_real" in like.aux:
            y.aux = like.aux
        return y

    def act(y, a):
        return torch.relu(y) if a == _lib.ACT_RELU else (torch.tanh(y) if a == _lib.ACT_TANH else y)

    def conv(x, channels, opt, kernel=4, stride=2, pad=0, dilation=1, use_bias=True, scope='conv_0', _residual=None, _act=_lib.ACT_NONE,
             _pad_cout_to=0, _pre_relu=False):
        sync()
        xin = torch.relu(x.data) if _pre_relu else x.data       # conv(relu(x)): the activation resblock_down fuses into its first conv
        y = R.conv(rs, xin, channels, opt, kernel, stride, pad, dilation, use_bias, scope)
        return T(act(y if _residual is None else y + _residual.data, _act))

    def deconv(x, channels, opt, kernel=4, stride=2, padding='SAME', use_bias=True, scope='deconv_0', _residual=None):
        sync()
        y = R.deconv(rs, x.data, channels, opt, kernel, stride, padding, use_bias, scope)
        return T(y if _residual is None else y + _residual.data)

    def fully_connected(x, units, opt, use_bias=True, lrmul=1.0, scope='fully_0'):
        sync()
        return T(R.fully_connected(rs, x.data, units, opt, use_bias, lrmul, scope))

    def condition_batch_norm(x, z, opt={}, scope='batch_norm', _act=_lib.ACT_NONE):
        sync()
        return T(act(R.condition_batch_norm(rs, x.data, z.data, opt, scope), _act))

    def batch_norm(x, opt={}, scope='batch_norm', _act=_lib.ACT_NONE):
        sync()
        return T(act(R.batch_norm(rs, x.data, opt, scope), _act))

    def self_attention_2(x, channels, opt, scope='self_attention'):
        sync()
        return T(R.self_attention_2(rs, x.data, channels, opt, scope))

    def prelu(x, scope=None, init_val=0.0):
        sync()
        return T(R.make_prelu(rs)(x.data, scope, init_val))

    leaves = {
        "conv": conv, "deconv": deconv, "fully_connected": fully_connected, "condition_batch_norm": condition_batch_norm,
        "batch_norm": batch_norm, "self_attention_2": self_attention_2, "prelu": prelu,
        "relu": lambda x: T(torch.relu(x.data)), "lrelu": lambda x, alpha=0.2: T(R.lrelu(x.data, alpha)),
        "avg_pooling": lambda x: T(R.avg_pooling(x.data)), "max_pooling": lambda x: T(R.max_pooling(x.data)),
        "up_sample": lambda x, scale_factor=2: T(R.up_sample(x.data, scale_factor)),
        "depth_to_space": lambda x, block_size: T(R.depth_to_space(x.data, block_size)),
        "add": lambda a, b: T(a.data + b.data),
        "global_sum_pooling": lambda x, _relu_first=False: T((torch.relu(x.data) if _relu_first else x.data).sum(dim=(1, 2))),
        "cast": lambda x, dtype: x, "reshape": lambda x, shape: T(x.data.reshape([int(s) for s in shape])),
    }
    for name, fn in leaves.items():
        monkeypatch.setattr(ops, name, fn)
        if hasattr(model, name):
            monkeypatch.setattr(model, name, fn)
    yield rs, st
    tf_shim.set_store(tf_shim.VariableStore())


@pytest.mark.parametrize("name", ["net64", "net128da", "net256", "net512", "net64cli", "net64variants"])
def test_model_and_composite_ops_wire_the_reference_networks(leaf_stand_ins, name):
    from biggan_tensorflow_b200 import model
    from biggan_tensorflow_b200.tensor import Tensor
    rs, st = leaf_stand_ins
    kw = C.NETS[name] if name in C.NETS else C.NEXT_NETS[name]
    g = np.load(os.path.join(HERE, "golden", "ref_%s.npz" % name))
    cfg = model.Config(**kw)
    b = M.synthetic_batch(M.Config(**kw), B=kw["batch_size"])
    f64 = lambda a: torch.as_tensor(a, dtype=torch.float64)
    rnd = [None] * 3
    if kw.get("da_policy"):
        rnd = [R.rnd_to_torch({k.split("/")[1]: g[k] for k in g.files if k.startswith("rnd%d/" % i)}) for i in range(3)]
    # D phase of build_model (BigGAN.py:806-807, :857-883): D(aug(real)), G(z), D(aug(G(z)))
    real = R.DiffAugment(f64(b["real"]), policy=kw.get("da_policy", ""), rnd=rnd[0])
    d_real = model.discriminator(cfg, Tensor(real))
    fake = model.generator(cfg, f64(b["z_d"]), f64(b["labels"]))
    d_fake = model.discriminator(cfg, Tensor(R.DiffAugment(fake.data, policy=kw.get("da_policy", ""), rnd=rnd[1])))
    C.check(g, "d_fake_images", fake.data.detach().numpy(), 2e-9)
    C.check(g, "d_real_logits", d_real["real"].data.detach().numpy(), 2e-9)
    C.check(g, "d_fake_logits", d_fake["real"].data.detach().numpy(), 2e-9)
    # every variable of the reference graph exists under the same name with the same shape, and nothing else
    ref_shapes = {str(n): tuple(int(s) for s in str(sh).split(",") if s) for n, sh in zip(g["all_names"], g["all_shapes"])}
    ours = {k: tuple(v.shape) for k, v in rs.vars.items()}
    ours.update({k: tuple(v.data.shape) for k, v in st.vars.items()})
    assert ours == ref_shapes
    # the regulariser is attached to exactly the generator weights the reference regularises (BigGAN.py:257-271)
    if kw.get("g_regularization", "none") != "none":
        assert len(rs.regularizers) == int(g["n_regularization_losses"])
        assert all(n.startswith("generator/") and n.endswith("/kernel") for n in rs.regularizers)
    else:
        assert not rs.regularizers
