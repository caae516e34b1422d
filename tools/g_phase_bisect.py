"""Diagnostic: where does the fp32-mode error of the G-phase gradients enter?  Compares, on net64:
 (1) G forward (fake image), (2) d loss / d fake (D's input-gradient chain) given the SAME image, against the float64 oracle
 running on the CUDA path's own state."""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from golden import ref_cases as C                                   # noqa: E402
from oracle import ref_model as M, ref_ops as R                     # noqa: E402
from biggan_tensorflow_b200.model import Config                     # noqa: E402
from biggan_tensorflow_b200.step import Trainer                     # noqa: E402

name = sys.argv[1] if len(sys.argv) > 1 else "net64"
kw = C.NETS[name]
cfg = M.Config(**kw)
b = M.synthetic_batch(cfg, B=kw["batch_size"])
dev = lambda a: torch.from_numpy(a).cuda()
tr = Trainer(Config(**kw), compute_dtype=torch.float32)
tr.build()
tr.d_update(dev(b["real"]), dev(b["labels"]), dev(b["z_d"]), dev(b["labels"]), apply=False)
state = tr.store.export()
g_loss, aux = tr.g_update(dev(b["z_g"]), dev(b["labels"]), apply=False)
fake = aux["fake"]
img = fake.numpy()[..., :3]
dimg = fake.grad.float().cpu().numpy()[..., :3]


def rel(a, r):
    return float(np.abs(a - r).max() / np.abs(r).max()), float(np.linalg.norm(a - r) / np.linalg.norm(r))


for dt in (torch.float64, torch.float32):
    rs = R.RefStore(dt, params=state)
    rs.begin_phase()
    M._detach_all(rs)
    cls_t = torch.as_tensor(b["labels"], dtype=dt)
    fake_ref = M.generator(rs, cfg, torch.as_tensor(b["z_g"], dtype=dt), cls_t)
    leaf = torch.tensor(img, dtype=dt, requires_grad=True)
    out = M.discriminator(rs, cfg, leaf)
    loss = R.generator_loss(cfg.gan_type, fake=out["real"], real=None) + cfg.g_cls_loss_weight * R.cls_loss_logistic(cls_t, out["cls"])
    (gi,) = torch.autograd.grad(loss, [leaf])
    print(dt, "G forward image: max-abs/max %.2e relL2 %.2e" % rel(img, fake_ref.detach().double().numpy()))
    print(dt, "dL/d image given the same image: max-abs/max %.2e relL2 %.2e" % rel(dimg, gi.double().numpy()))
    e = np.abs(dimg - gi.double().numpy()).max(axis=(0, 3)) / np.abs(gi.double().numpy()).max()
    np.set_printoptions(linewidth=250, precision=1)
    print("error map (max over batch, channel; rows/cols subsampled by 4), units of 1e-6:")
    print((e[::4, ::4] * 1e6).astype(int))

# (3) all G gradients against the float64 oracle started from the CUDA path's own post-D state
gg = tr.grads("generator")
rs = R.RefStore(torch.float64, params=state)
_, ref_gg, _ = M.g_update(rs, cfg, b["z_g"], b["labels"], apply=False)
worst = sorted(((rel(gg[n], ref_gg[n].numpy())[0], n) for n in gg if not n.endswith("f_conv/bias")), reverse=True)[:6]
print("G grads vs float64 oracle on the SAME post-D state, worst max-abs/max:", [("%.2e" % e, n) for e, n in worst])
# (4) the CUDA path's post-D state against the golden post-D state
g = np.load(os.path.join(ROOT, "tests", "golden", "ref_%s.npz" % name))
rows = []
for k in sorted(set(f.split("#")[0] for f in g.files if f.startswith("state_after_D/"))):
    n = k[len("state_after_D/"):]
    v = np.asarray(state[n], dtype=np.float64)
    if k in g.files:
        r = g[k].astype(np.float64)
        v = v.reshape(r.shape)
    else:
        r = g[k + "#sub"].astype(np.float64)
        v = v.ravel()[::int(g[k + "#stat"][1])]
    rows.append((float(np.abs(v - r).max() / max(np.abs(r).max(), 1e-6)), n))
rows.sort(reverse=True)
print("post-D state vs golden, worst:", [("%.2e" % e, n) for e, n in rows[:8]])

# (5) G backward alone: the oracle's generator back-propagates the CUDA path's OWN dL/d image (removes the kinks of D --
# ReLU masks, max-pool ties -- from the comparison)
rs = R.RefStore(torch.float64, params=state)
rs.begin_phase()
M._detach_all(rs)
cls_t = torch.as_tensor(b["labels"], dtype=torch.float64)
fake_ref = M.generator(rs, cfg, torch.as_tensor(b["z_g"], dtype=torch.float64), cls_t)
g_vars = rs.trainables("generator")
names = list(g_vars)
grads = torch.autograd.grad(fake_ref, [g_vars[n] for n in names], grad_outputs=torch.as_tensor(dimg, dtype=torch.float64), allow_unused=True)
worst = sorted(((rel(gg[n], gr.numpy())[0], n) for n, gr in zip(names, grads) if gr is not None and float(gr.abs().max()) > 0), reverse=True)[:6]
print("G grads vs float64 oracle G-backward seeded with the CUDA dL/d image, worst max-abs/max:", [("%.2e" % e, n) for e, n in worst])
