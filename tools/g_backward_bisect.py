"""Diagnostic: activation GRADIENTS of the G phase, layer by layer, CUDA fp32 path vs the float64 oracle started from the
CUDA path's own post-D state.  Layer outputs are matched by (parent scope, function, call index)."""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from golden import ref_cases as C                                   # noqa: E402
from oracle import ref_model as M, ref_ops as R                     # noqa: E402
from biggan_tensorflow_b200 import ops, tf_shim                     # noqa: E402
from biggan_tensorflow_b200.model import Config                     # noqa: E402
from biggan_tensorflow_b200.step import Trainer                     # noqa: E402

name = sys.argv[1] if len(sys.argv) > 1 else "net64"
kw = C.NETS[name]
cfg = M.Config(**kw)
b = M.synthetic_batch(cfg, B=kw["batch_size"])
dev = lambda a: torch.from_numpy(a).cuda()
FNS = ["conv", "deconv", "cond_bn", "bn", "self_attention_2", "fully_connected"]
prod, orac = {}, {}


def wrap_prod(fn_name):
    f = getattr(ops, fn_name)

    def w(*a, **k):
        parent = tf_shim.current_scope_name()
        y = f(*a, **k)
        key = (parent, fn_name, k.get("scope", ""))
        key = key + (sum(1 for q in prod if q[:3] == key),)
        prod[key] = y
        return y
    setattr(ops, fn_name, w)


def wrap_orac(fn_name):
    f = getattr(R, fn_name)

    def w(store, *a, **k):
        parent = store.full_name()
        y = f(store, *a, **k)
        key = (parent, fn_name, k.get("scope", ""))
        key = key + (sum(1 for q in orac if q[:3] == key),)
        if y.requires_grad:
            y.retain_grad()
        orac[key] = y
        return y
    setattr(R, fn_name, w)


tr = Trainer(Config(**kw), compute_dtype=torch.float32)
tr.build()
tr.d_update(dev(b["real"]), dev(b["labels"]), dev(b["z_d"]), dev(b["labels"]), apply=False)
state = tr.store.export()
for n in FNS:
    wrap_prod(n)
    wrap_orac(n)
import biggan_tensorflow_b200.model as PM                           # noqa: E402
for n in FNS:                                                       # model.py imported the names directly
    if hasattr(PM, n):
        setattr(PM, n, getattr(ops, n))
g_loss, aux = tr.g_update(dev(b["z_g"]), dev(b["labels"]), apply=False)
torch.cuda.synchronize()
rs = R.RefStore(torch.float64, params=state)
rs.begin_phase()
M._detach_all(rs)
cls_t = torch.as_tensor(b["labels"], dtype=torch.float64)
fake = M.generator(rs, cfg, torch.as_tensor(b["z_g"], dtype=torch.float64), cls_t)
out = M.discriminator(rs, cfg, fake)
loss = R.generator_loss(cfg.gan_type, fake=out["real"], real=None) + cfg.g_cls_loss_weight * R.cls_loss_logistic(cls_t, out["cls"])
loss.backward()
print("keys only in product:", sorted(set(prod) - set(orac))[:5], " only in oracle:", sorted(set(orac) - set(prod))[:5])
for key in orac:                 # forward order of the oracle
    if key not in prod:
        continue
    yo, yp = orac[key], prod[key]
    v = yp.data.float().cpu().numpy().astype(np.float64)
    r = yo.detach().numpy()
    if v.size != r.size:
        v = v[..., :r.shape[-1]]
    fe = np.abs(v.reshape(r.shape) - r).max() / max(np.abs(r).max(), 1e-30)
    ge = float("nan")
    if yo.grad is not None and yp.grad is not None:
        gv = yp.grad.float().cpu().numpy().astype(np.float64)
        gr = yo.grad.numpy()
        if gv.size != gr.size:
            gv = gv[..., :gr.shape[-1]]
        gv = gv.reshape(gr.shape)
        if key[1] in ("bn", "cond_bn"):                      # product output is post-ReLU: compare where the unit is active
            m = (r > 1e-4)
            fe = np.abs(v.reshape(r.shape) - r)[m].max() / max(np.abs(r).max(), 1e-30)
            ge = np.abs(gv - gr)[m].max() / max(np.abs(gr).max(), 1e-30)
        elif key[2] == "G_logit":                     # product output is post-tanh
            t = np.tanh(r)
            fe = np.abs(v.reshape(r.shape) - t).max()
            ge = np.abs(gv * (1 - t * t) - gr).max() / max(np.abs(gr).max(), 1e-30)
        else:
            ge = np.abs(gv - gr).max() / max(np.abs(gr).max(), 1e-30)
    print("%-58s %-18s fwd %.1e  grad %.1e  %s" % (key[0] + "/" + key[2], key[1], fe, ge, tuple(r.shape)))

key = [k for k in orac if k[1] == "bn"][0]
gr = orac[key].grad.numpy()
gv = prod[key].grad.float().cpu().numpy().astype(np.float64)[..., :gr.shape[-1]]
m = orac[key].detach().numpy() > 1e-4
e = np.abs(gv - gr) * m / np.abs(gr).max()
np.set_printoptions(linewidth=250)
print("final-BN output gradient error map (max over batch, channel; subsampled by 4), 1e-6 units")
print((e.max(axis=(0, 3))[::4, ::4] * 1e6).astype(int))
print("per batch sample:", e.max(axis=(1, 2, 3)), " per channel:", (e.max(axis=(0, 1, 2)) * 1e6).astype(int))
kl = [k for k in orac if k[2] == "G_logit"][0]
pre = orac[kl].detach().numpy()
print("pre-tanh |x| max %.2f, mean %.2f; fraction |x| > 4: %.4f" % (np.abs(pre).max(), np.abs(pre).mean(), (np.abs(pre) > 4).mean()))
