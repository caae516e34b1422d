"""One full G+D training iteration (D update then G update, BigGAN.py:1061-1084) on BigGAN-32
(BASELINE config 0 shape, reduced batch) -- CUDA path vs the CPU oracle on identical weights/inputs."""
import numpy as np
import pytest
import torch

from helpers import TOL, assert_close
from oracle import ref_model as M
from oracle import ref_ops as R

pytestmark = pytest.mark.gpu


def _run(dtype, da_policy, B=8, n_labels=10, img=32, ch=32):
    from biggan_tensorflow_b200 import DiffAugment_tf as D
    from biggan_tensorflow_b200.model import Config
    from biggan_tensorflow_b200.step import Trainer
    kw = dict(img_size=img, ch=ch, batch_size=B, n_labels=n_labels, da_policy=da_policy, sa_gamma_init=0.1)
    cfg, rcfg = Config(**kw), M.Config(**kw)
    batch = M.synthetic_batch(rcfg, B=B)
    rng = np.random.Generator(np.random.PCG64(5))
    rnds = [R.draw_diffaug_randoms(rng, B, img, img) for _ in range(3)]

    tr = Trainer(cfg, compute_dtype=dtype)
    tr.build()
    dev = lambda a: torch.from_numpy(a).cuda()
    real, labels, z_d, z_g = dev(batch["real"]), dev(batch["labels"]), dev(batch["z_d"]), dev(batch["z_g"])
    if da_policy:
        D.push_randoms(rnds[0]); D.push_randoms(rnds[1]); D.push_randoms(rnds[2])
    d_loss, _ = tr.d_update(real, labels, z_d, labels)
    d_grads = tr.grads("discriminator")
    g_loss, _ = tr.g_update(z_g, labels)
    g_grads = tr.grads("generator")
    prod = tr.store.export()
    ema = tr.opt["generator"]["ema"].cpu().numpy()

    rs = R.RefStore(torch.float64)
    od = M.AdamState(rcfg.d_lr, rcfg.beta1, rcfg.beta2)
    og = M.AdamState(rcfg.g_lr, rcfg.beta1, rcfg.beta2, rcfg.moving_decay)
    t = [R.rnd_to_torch(r) for r in rnds]
    rd_loss, rd_grads, _ = M.d_update(rs, rcfg, batch["real"], batch["labels"], batch["z_d"], batch["labels"],
                                      rnd_real=t[0], rnd_fake=t[1], opt_state=od)
    rg_loss, rg_grads, _ = M.g_update(rs, rcfg, batch["z_g"], batch["labels"], rnd_fake=t[2], opt_state=og)
    return dict(tr=tr, d_loss=float(d_loss.cpu()), g_loss=float(g_loss.cpu()), d_grads=d_grads, g_grads=g_grads,
                prod=prod, ema=ema, rs=rs, rd_loss=rd_loss, rg_loss=rg_loss, rd_grads=rd_grads, rg_grads=rg_grads, og=og)


@pytest.mark.parametrize("da_policy", ["", "color,translation,cutout"])
@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
def test_train_step_matches_oracle(dtype, da_policy):
    r = _run(dtype, da_policy)
    tol = TOL[dtype]
    assert set(r["prod"].keys()) == set(r["rs"].vars.keys())
    assert abs(r["d_loss"] - r["rd_loss"]) <= tol * max(abs(r["rd_loss"]), 1.0)
    assert abs(r["g_loss"] - r["rg_loss"]) <= tol * max(abs(r["rg_loss"]), 1.0)
    # Whole-network gradients compound the per-op error.  fp32: 3e-4 max-abs.  bf16: relative L2, calibrated
    # against the bf16-STORAGE emulation of the oracle itself (tests/test_bf16_conditioning.py): D gradients
    # (one D backward) 8e-2; G gradients (back through all of D and all of G, ~60 bf16-stored tensors;
    # the emulation alone moves generator/dense1 by 1.3e-1) 2.5e-1.
    for which, ref, got in (("D", r["rd_grads"], r["d_grads"]), ("G", r["rg_grads"], r["g_grads"])):
        gscale = max(float(g.abs().max()) for g in ref.values())
        for name, g in ref.items():
            if name.endswith("f_conv/bias"):      # analytically zero (softmax shift invariance)
                assert np.abs(got[name]).max() <= 1e-3 * gscale, name
                continue
            if dtype == torch.float32:
                vtol = 3e-4
            else:
                vtol = 8e-2 if which == "D" else 2.5e-1
                if "f_conv/kernel" in name or "h_conv/kernel" in name or "g_conv" in name:
                    vtol = max(vtol, 1.5e-1)
            assert_close(got[name], g.numpy(), vtol, which + " grad " + name)
    # state after the iteration: SN u vectors, BN running stats
    for name, v in r["rs"].vars.items():
        if not r["rs"].trainable.get(name, False):
            assert_close(r["prod"][name], v.detach().numpy(), max(tol, 1e-4) * 2, "state " + name)
    # EMA shadows exist for every generator trainable
    assert np.isfinite(r["ema"]).all()


def test_parameters_after_update_fp32():
    """Adam(+EMA) applied: parameters after one iteration match the oracle.  With beta1=0 the first Adam
    step is lr*sign(g), so only entries whose gradient is clearly non-zero are compared tightly."""
    r = _run(torch.float32, "")
    bad = 0
    total = 0
    for name, v in r["rs"].vars.items():
        if not r["rs"].trainable.get(name, False):
            continue
        ref = v.detach().numpy()
        got = r["prod"][name]
        g = (r["rd_grads"].get(name) if name in r["rd_grads"] else r["rg_grads"][name]).numpy()
        sure = np.abs(g) > 1e-3 * np.abs(g).max()
        total += int(sure.sum())
        bad += int((np.abs(got - ref)[sure] > 1e-6 + 1e-4 * np.abs(ref[sure])).sum())
    assert total > 1000 and bad <= total * 1e-3, (bad, total)
    ar = r["tr"].store.arenas["generator"]
    for v, o in zip(ar["vars"], ar["offsets"]):
        n = v.data.numel()
        sh = r["og"].shadow[v.name].numpy().reshape(-1)
        assert np.abs(r["ema"][o:o + n] - sh).max() <= 1e-5 * max(np.abs(sh).max(), 1e-3) + 1e-7, v.name
