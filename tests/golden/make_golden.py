#!/usr/bin/env python
"""Generates tests/golden/oracle_step_cfg0_b4.npz.

ORACLE-generated (the reference-generated vectors are tests/golden/ref_*.npz): BigGAN-32 is an extrapolated
shape the reference's generator() does not accept (BigGAN.py:290-294), so these vectors come from the CPU oracle
(oracle/, float64, itself pinned to the reference's source by tests/test_reference_golden.py) -- they freeze the oracle's
behaviour (a regression guard: `tests/test_golden.py` fails if the oracle drifts) and give the GPU tests a fixed target
that does not depend on running the oracle.  Regenerate with:  python tests/golden/make_golden.py
"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import ref_model as M  # noqa: E402
from oracle import ref_ops as R  # noqa: E402

KW = dict(img_size=32, ch=32, batch_size=4, n_labels=10, da_policy="", sa_gamma_init=0.1)


def oracle_step():
    cfg = M.Config(**KW)
    b = M.synthetic_batch(cfg, B=4)
    rs = R.RefStore(torch.float64)
    od = M.AdamState(cfg.d_lr, cfg.beta1, cfg.beta2)
    og = M.AdamState(cfg.g_lr, cfg.beta1, cfg.beta2, cfg.moving_decay)
    d_loss, d_grads, _ = M.d_update(rs, cfg, b["real"], b["labels"], b["z_d"], b["labels"], opt_state=od)
    g_loss, g_grads, _ = M.g_update(rs, cfg, b["z_g"], b["labels"], opt_state=og)
    out = {"d_loss": np.float64(d_loss), "g_loss": np.float64(g_loss)}
    # a few small, well-conditioned tensors: gradient norms of every variable + selected full gradients + updated state
    for which, grads in (("D", d_grads), ("G", g_grads)):
        names = sorted(grads)
        out[which + "_names"] = np.array(names)
        out[which + "_grad_norms"] = np.array([float(grads[n].norm()) for n in names])
    out["grad_D_logit_kernel"] = d_grads["discriminator/D_logit/kernel"].numpy()
    out["grad_G_logit_kernel"] = g_grads["generator/G_logit/kernel"].numpy()
    out["u_D_logit"] = rs.vars["discriminator/D_logit/u"].detach().numpy()
    out["u_G_dense2"] = rs.vars["generator/dense2/u"].detach().numpy()
    return out


if __name__ == "__main__":
    out = oracle_step()
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "oracle_step_cfg0_b4.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, {k: getattr(v, "shape", ()) for k, v in out.items()})
