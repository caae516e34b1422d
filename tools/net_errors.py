"""Diagnostic: per-variable error of the CUDA path's whole-network gradients against the reference-generated golden
vectors (tests/golden/ref_<net>.npz).   python tools/net_errors.py net64 [f32|bf16]"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from golden import ref_cases as C                                   # noqa: E402
from oracle import ref_model as M                                   # noqa: E402

name = sys.argv[1] if len(sys.argv) > 1 else "net64"
dtype = torch.bfloat16 if (len(sys.argv) > 2 and sys.argv[2] == "bf16") else torch.float32
from biggan_tensorflow_b200 import DiffAugment_tf as D              # noqa: E402
from biggan_tensorflow_b200.model import Config                     # noqa: E402
from biggan_tensorflow_b200.step import Trainer                     # noqa: E402
g = np.load(os.path.join(ROOT, "tests", "golden", "ref_%s.npz" % name))
kw = C.NETS[name] if name in C.NETS else C.NEXT_NETS[name]
b = M.synthetic_batch(M.Config(**kw), B=kw["batch_size"])
dev = lambda a: torch.from_numpy(a).cuda()
tr = Trainer(Config(**kw), compute_dtype=dtype)
tr.build()
if kw.get("da_policy"):
    for i in range(3):
        D.push_randoms({k.split("/")[1]: g[k] for k in g.files if k.startswith("rnd%d/" % i)})
d_loss, _ = tr.d_update(dev(b["real"]), dev(b["labels"]), dev(b["z_d"]), dev(b["labels"]), apply=False)
gd = tr.grads("discriminator")
g_loss, _ = tr.g_update(dev(b["z_g"]), dev(b["labels"]), apply=False)
gg = tr.grads("generator")
print("d_loss %.7f ref %.7f   g_loss %.7f ref %.7f" % (float(d_loss.cpu()), float(g["d_loss"]), float(g_loss.cpu()), float(g["g_loss"])))
for which, grads in (("D", gd), ("G", gg)):
    names = [str(n) for n in g[which + "_names"]]
    ref = g[which + "_grad_norms"]
    for n, rn in zip(names, ref):
        v = np.asarray(grads[n], dtype=np.float64)
        k = which + "_grad/" + n
        e = float("nan")
        if k in g.files:
            r = g[k].astype(np.float64)
            e = np.abs(v.reshape(r.shape) - r).max() / max(np.abs(r).max(), 1e-6)
        elif k + "#sub" in g.files:
            stride = int(g[k + "#stat"][1])
            r = g[k + "#sub"].astype(np.float64)
            e = np.abs(v.ravel()[::stride] - r).max() / max(np.abs(r).max(), 1e-6)
        nrm = float(np.linalg.norm(v))
        print("%s %-62s norm rel err %.2e   max-abs rel err %.2e   |ref| %.3e" % (which, n, abs(nrm - rn) / max(rn, 1e-30), e, rn))
