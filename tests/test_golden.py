"""Golden vectors of one BigGAN-32 training iteration (batch 4): `tests/golden/oracle_step_cfg0_b4.npz`, written by
`tests/golden/make_golden.py` FROM THE ORACLE (the reference's generator() has no 32x32 table; the reference-generated
vectors live in tests/test_reference_golden.py).
CPU: the oracle still reproduces them bit-for-bit-ish (drift guard).  GPU: the CUDA path reproduces them through the C ABI."""
import importlib.util
import os

import numpy as np
import pytest
import torch

from helpers import assert_close

HERE = os.path.dirname(os.path.abspath(__file__))
GOLD = os.path.join(HERE, "golden", "oracle_step_cfg0_b4.npz")


def _maker():
    spec = importlib.util.spec_from_file_location("make_golden", os.path.join(HERE, "golden", "make_golden.py"))
    m = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(m)
    return m


def test_oracle_reproduces_golden_vectors():
    g = np.load(GOLD)
    out = _maker().oracle_step()
    assert abs(float(out["d_loss"]) - float(g["d_loss"])) < 1e-9 and abs(float(out["g_loss"]) - float(g["g_loss"])) < 1e-9
    for k in ("D_grad_norms", "G_grad_norms", "grad_D_logit_kernel", "grad_G_logit_kernel", "u_D_logit", "u_G_dense2"):
        np.testing.assert_allclose(out[k], g[k], rtol=1e-7, atol=1e-12, err_msg=k)
    assert list(out["D_names"]) == list(g["D_names"]) and list(out["G_names"]) == list(g["G_names"])


@pytest.mark.gpu
def test_cuda_step_matches_golden_vectors():
    from biggan_tensorflow_b200.model import Config
    from biggan_tensorflow_b200.step import Trainer
    from oracle import ref_model as M
    m = _maker()
    g = np.load(GOLD)
    cfg = Config(**m.KW)
    b = M.synthetic_batch(M.Config(**m.KW), B=4)
    tr = Trainer(cfg, compute_dtype=torch.float32)
    tr.build()
    dev = lambda a: torch.from_numpy(a).cuda()
    d_loss, _ = tr.d_update(dev(b["real"]), dev(b["labels"]), dev(b["z_d"]), dev(b["labels"]))
    dg = tr.grads("discriminator")
    g_loss, _ = tr.g_update(dev(b["z_g"]), dev(b["labels"]))
    gg = tr.grads("generator")
    assert abs(float(d_loss.cpu()) - float(g["d_loss"])) <= 1e-4 * max(1.0, abs(float(g["d_loss"])))
    assert abs(float(g_loss.cpu()) - float(g["g_loss"])) <= 1e-4 * max(1.0, abs(float(g["g_loss"])))
    for which, grads in (("D", dg), ("G", gg)):
        norms = np.array([float(np.linalg.norm(grads[n].astype(np.float64))) for n in g[which + "_names"]])
        ref = g[which + "_grad_norms"]
        ok = ref > 1e-3 * ref.max()                     # analytically-zero gradients (f_conv bias) hold only noise
        np.testing.assert_allclose(norms[ok], ref[ok], rtol=3e-4, err_msg=which + " gradient norms")
    assert_close(dg["discriminator/D_logit/kernel"], g["grad_D_logit_kernel"], 3e-4, "d/d D_logit/kernel")
    assert_close(gg["generator/G_logit/kernel"], g["grad_G_logit_kernel"], 3e-4, "d/d G_logit/kernel")
    st = tr.store.export()
    assert_close(st["discriminator/D_logit/u"], g["u_D_logit"], 2e-4, "u D_logit")
    assert_close(st["generator/dense2/u"], g["u_G_dense2"], 2e-4, "u dense2")
