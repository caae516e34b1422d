"""GPU half of tests/test_reference_golden.py: the CUDA path (through the C ABI) against the golden vectors produced by
the REFERENCE'S OWN SOURCE (`tests/golden/ref_*.npz`, see `oracle/make_reference_golden.py`).  Tolerances are
north_star's: 1e-4 relative in fp32 mode, 2e-2 in bf16 mode (with the per-case widening factors tests/test_ops_gpu.py
uses against the oracle on the same shapes); DiffAugment index work bit-exact.  The oracle is not involved here: the
expected values come straight from the committed fixtures."""
import os

import numpy as np
import pytest
import torch

from test_reference_golden import C, M, check_net64, gold

GPU_TOL = {torch.float32: 1e-4, torch.bfloat16: 2e-2}     # north_star
# per-case widening factors, identical to the ones tests/test_ops_gpu.py uses against the oracle on the same shapes
TOL_SCALE = {"cbn_relu_4x4x64": (2.0, 2.0), "cbn_relu_8x8x32": (2.0, 2.0), "bn_train": (2.0, 2.0),
             "sa2_8x8x32": (2.0, 4.0), "sa2_16x16x64": (2.0, 4.0), "resblock_up_condition": (4.0, 7.5),
             "resblock_down_rgb": (3.0, 3.0), "resblock_down_32": (3.0, 3.0), "resblock": (3.0, 3.0),
             # SURVEY 8(f) cases: the same factors as their canonical counterparts
             "resblock_down_resize_conv1": (3.0, 3.0), "resblock_down_bn_in_d": (3.0, 3.0), "resblock_bn_in_d_prelu": (3.0, 3.0),
             "resblock_up_condition_prelu": (4.0, 7.5), "resblock_up_condition_lrelu_subpixel3_conv5": (4.0, 7.5),
             "cbrn_fresh": (2.0, 2.0), "cbrn_warm": (2.0, 2.0), "cbrn_shared": (2.0, 2.0), "cbrn_inference": (2.0, 2.0)}
FP32_ONLY = {"fc_sn", "fc_dense", "cbn_inference", "sa_full_4x4x32"}     # as in tests/test_ops_gpu.py


def UNVERIFIED(f):
    """Round 1 gated these tests behind BG_RUN_UNVERIFIED=1 (written without GPU time).  They ran on B200 at the start of
    round 2 and are part of the normal `-m gpu` suite now; the marker is kept as a no-op to show which they were."""
    return f


def _acts(ops):
    return {"relu": ops.relu, "prelu": ops.prelu, "lrelu": lambda x: ops.lrelu(x, 0.2)}


def _product_call(case, ops, _lib, xs, opt):
    fn, kw = case["fn"], dict(case["kw"])
    if fn in ("prelu", "lrelu"):
        return getattr(ops, fn)(*xs.values())
    if fn == "cond_bn" and case["relu_after"]:
        return ops.cond_bn(xs["x"], xs["z"], opt, _act=_lib.ACT_RELU)        # CBN + ReLU fused, as the resblocks call it
    y = getattr(ops, fn)(*xs.values(), **kw, opt=opt)
    assert not case["relu_after"]
    return y


@pytest.mark.gpu
@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16], ids=["fp32", "bf16"])
@pytest.mark.parametrize("case", C.OP_CASES, ids=[c["name"] for c in C.OP_CASES])
def test_cuda_op_matches_reference_source(case, dtype):
    if dtype == torch.bfloat16 and case["name"] in FP32_ONLY:
        pytest.skip("fp32-only case (same split as tests/test_ops_gpu.py)")
    _cuda_op_case(case, dtype, "ref_ops.npz")


@pytest.mark.gpu
@UNVERIFIED
@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16], ids=["fp32", "bf16"])
@pytest.mark.parametrize("case", C.NEXT_OP_CASES, ids=[c["name"] for c in C.NEXT_OP_CASES])
def test_cuda_next_row_op_matches_reference_source(case, dtype):
    """prelu / lrelu, up- and down-sampling and g_conv variants, bn_in_d, conditional batch renorm."""
    _cuda_op_case(case, dtype, "ref_ops_next.npz")


def _cuda_op_case(case, dtype, fname):
    from helpers import dev_tensor, new_product_store, seed_backward
    from biggan_tensorflow_b200 import _lib, ops
    g = gold(fname)
    p = case["name"]
    bf = dtype == torch.bfloat16
    tol = GPU_TOL[dtype] * TOL_SCALE.get(p, (1.0, 1.0))[1 if bf else 0]
    st = new_product_store(dtype)
    if case["var_init"]:
        st.load({k: np.asarray(v, np.float32) for k, v in case["var_init"].items()})
    opt = C.make_opt(_acts(ops), case["opt"])
    xs = {k: dev_tensor(v, torch.float32 if k in C.FP32_INPUTS else dtype) for k, v in C.case_inputs(case).items()}
    y = _product_call(case, ops, _lib, xs, opt)
    seed_backward(y, C.case_dy(case, tuple(y.shape)))
    C.check(g, p + "/out", y.numpy(), tol, l2=bf)
    for k in xs:
        key = p + "/gin/" + k
        if key in g.files or key + "#stat" in g.files:
            assert xs[k].grad is not None, "no gradient reached input " + k
            C.check(g, key, xs[k].grad.float().cpu().numpy(), tol, l2=bf)
    ref_names = sorted(k.split("/varsum/")[1] for k in g.files if k.startswith(p + "/varsum/"))
    assert sorted(st.vars.keys()) == ref_names
    gkeys = sorted(set(k.split("#")[0] for k in g.files if k.startswith(p + "/gvar/")))
    gscale = max([C.golden_absmax(g, k) for k in gkeys] + [0.0])
    for key in gkeys:
        n = key.split("/gvar/")[1]
        if n.endswith("f_conv/bias") or C.golden_absmax(g, key) <= 1e-10 * gscale:
            continue                                   # analytically zero (only rounding noise on both sides)
        pv = st.vars[n]
        assert pv.grad is not None, "no gradient reached variable " + n
        vtol = tol
        if bf and ("f_conv/kernel" in n or "h_conv/kernel" in n or n.endswith("g_conv/bias")):
            vtol = 1.5e-1                              # max-pooled attention branches under bf16: tests/test_bf16_conditioning.py
        C.check(g, key, pv.grad.float().cpu().numpy(), vtol, l2=bf)
    for key in (k for k in g.files if k.startswith(p + "/state/")):
        n = key.split("/state/")[1]
        C.check(g, key, st.vars[n].data.float().cpu().numpy(), max(GPU_TOL[dtype], 1e-4), l2=bf)


@pytest.mark.gpu
@pytest.mark.parametrize("K,Cc", C.SN_CASES)
def test_cuda_spectral_norm_two_runs_match_reference_source(K, Cc):
    from helpers import new_product_store, seed_backward
    from biggan_tensorflow_b200 import ops
    from biggan_tensorflow_b200.tensor import clear_tape
    from biggan_tensorflow_b200.tf_shim import get_variable, variable_scope
    g = gold("ref_ops.npz")
    st = new_product_store(torch.float32)
    for run in (1, 2):
        st.begin_phase()
        with variable_scope("t"):
            w = get_variable("kernel", [K, Cc], initializer=ops.weight_init)
            wn = ops.spectral_norm(w)
        p = "sn_%dx%d/run%d" % (K, Cc, run)
        if run == 1:                                   # gradient checked on the first run (it accumulates on the product side)
            dy = np.random.default_rng(1000 + run).standard_normal((K, Cc)).astype(np.float32)
            seed_backward(wn, dy)
            C.check(g, p + "/gvar", w.grad.float().cpu().numpy(), 1e-4)
        clear_tape()
        C.check(g, p + "/out", wn.numpy(), 1e-4)
        C.check(g, p + "/u", st.vars["t/u"].data.float().cpu().numpy(), 1e-4)


@pytest.mark.gpu
@pytest.mark.parametrize("name,policy,shape", C.DIFFAUG_CASES, ids=[c[0] for c in C.DIFFAUG_CASES])
def test_cuda_diffaugment_matches_reference_source(name, policy, shape):
    """Index work (translation, cutout) bit-exact; colour arithmetic within 1e-5 (fp32)."""
    from helpers import dev_tensor, new_product_store, seed_backward
    from biggan_tensorflow_b200 import DiffAugment_tf as D
    g = gold("ref_diffaug.npz")
    new_product_store(torch.float32)
    x, dy = C.diffaug_inputs(shape)
    D.push_randoms({k.split("/rnd/")[1]: g[k] for k in g.files if k.startswith(name + "/rnd/")})
    tx = dev_tensor(x)
    y = D.DiffAugment(tx, policy=policy)
    seed_backward(y, dy)
    exact = "color" not in policy
    C.check(g, name + "/out", y.numpy(), 0.0 if exact else 1e-5)
    C.check(g, name + "/gx", tx.grad.float().cpu().numpy(), 1e-6 if exact else 1e-5)


def _to_cuda(a):
    return torch.from_numpy(a).cuda()


# fp32-mode tolerances of the whole-network checks (the error budget is measured):
#  * everything of the D phase, the G-phase losses and images: NET_TOL / NET_GTOL (north_star's 1e-4; 3e-4 on gradients);
#  * G-phase gradients are checked in two TIGHT halves split at the generated image, each against the float64 oracle
#    evaluated AT THE SAME POINT (the CUDA path's own state and image): dL/d image (D's input-gradient chain) and G's
#    backward seeded with that dL/d image.  Measured on B200: 7e-7 and 2.4e-6 (net64), 7e-7 and 6.5e-4 (net512, where
#    the float32 CPU run of the reference algorithm itself is 7.7e-4 off float64: BatchNorm sums over 2 x 512 x 512 terms);
#  * end to end against the golden vectors the same gradients are limited by the REFERENCE FUNCTION, not by arithmetic:
#    D(G(z)) is not differentiable at ReLU / max-pool ties, and a 2.5e-6 forward difference in the image moves one tie in
#    D, which changes dL/d image by 4e-3 of its maximum inside one receptive field of one sample (tools/g_backward_bisect.py)
#    and every G gradient by up to 2e-3 (net64) / 4e-3 (net512) of its maximum.  Hence KINK_GTOL for that comparison only.
NET_TOL, NET_GTOL, KINK_GTOL, KINK_NORM_TOL = 1e-4, 3e-4, 1e-2, 3e-3
SPLIT_TOL = {"net64": (1e-5, 1e-4), "net128da": (1e-5, 1e-4), "net256": (1e-5, 1e-3), "net512": (1e-5, 2e-3)}


def _rel_max(a, r):
    a, r = np.asarray(a, np.float64), np.asarray(r, np.float64)
    return float(np.abs(a.reshape(r.shape) - r).max() / max(float(np.abs(r).max()), 1e-30))


def _g_phase_split_check(name, kw, g, state_after_d, b, img, dimg, gg, tol_d, tol_g):
    """The two tight halves of the G-phase gradient check (see the tolerance note above).  float64 oracle, started from
    the CUDA path's own post-D state: (A) d loss / d image evaluated at the CUDA path's image; (B) the generator's backward
    seeded with the CUDA path's d loss / d image."""
    from oracle import ref_ops as R
    cfg = M.Config(**kw)
    dt = torch.float64
    rnd = None
    if kw.get("da_policy"):
        rnd = R.rnd_to_torch({k.split("/")[1]: g[k] for k in g.files if k.startswith("rnd2/")})
    rs = R.RefStore(dt, params=state_after_d)
    rs.begin_phase()
    M._detach_all(rs)
    cls_t = torch.as_tensor(b["labels"], dtype=dt)
    fake_ref = M.generator(rs, cfg, torch.as_tensor(b["z_g"], dtype=dt), cls_t)
    leaf = torch.tensor(np.asarray(img, np.float64), dtype=dt, requires_grad=True)
    out = M.discriminator(rs, cfg, R.DiffAugment(leaf, policy=cfg.da_policy, rnd=rnd))
    loss = R.generator_loss(cfg.gan_type, fake=out["real"], real=None, flood_level=cfg.g_flood)
    if cfg.acgan:
        loss = loss + cfg.g_cls_loss_weight * R.cls_loss_logistic(cls_t, out["cls"])
    (gi,) = torch.autograd.grad(loss, [leaf])
    e = _rel_max(dimg, gi.numpy())
    assert e <= tol_d, "%s: d loss / d image at the same image: error %.2e > %.0e" % (name, e, tol_d)
    g_vars = rs.trainables("generator")
    names = list(g_vars)
    grads = torch.autograd.grad(fake_ref, [g_vars[n] for n in names], grad_outputs=torch.as_tensor(np.asarray(dimg, np.float64)),
                                allow_unused=True)
    scale = max(float(gr.abs().max()) for gr in grads if gr is not None)
    for n, gr in zip(names, grads):
        if gr is None or n.endswith("f_conv/bias") or float(gr.abs().max()) <= 1e-10 * scale:
            continue                                    # analytically zero: only rounding noise on both sides
        e = _rel_max(gg[n], gr.numpy())
        assert e <= tol_g, "%s: G backward seeded with the same d loss / d image, %s: error %.2e > %.0e" % (name, n, e, tol_g)


def _cuda_net(name, dtype=torch.float32, stol=2e-4, dev=_to_cuda, tol=NET_TOL, gtol=NET_GTOL, l2=False, skip=(), split=True):
    from biggan_tensorflow_b200 import DiffAugment_tf as D
    from biggan_tensorflow_b200.model import Config
    from biggan_tensorflow_b200.step import Trainer
    g = gold("ref_%s.npz" % name)
    kw = C.NETS[name] if name in C.NETS else C.NEXT_NETS[name]
    b = M.synthetic_batch(M.Config(**kw), B=kw["batch_size"])
    tr = Trainer(Config(**kw), compute_dtype=dtype)
    tr.build()
    if kw.get("da_policy"):
        for i in range(3):
            D.push_randoms({k.split("/")[1]: g[k] for k in g.files if k.startswith("rnd%d/" % i)})
    host = lambda t: t.data.detach().float().cpu().numpy()
    d_loss, da = tr.d_update(dev(b["real"]), dev(b["labels"]), dev(b["z_d"]), dev(b["labels"]), apply=False)
    gd = tr.grads("discriminator")
    state_d = tr.store.export()
    d_aux = {"real_logits": host(da["real_logits"]), "fake_logits": host(da["fake_logits"]), "fake": host(da["fake"])[..., :3]}
    g_loss, ga = tr.g_update(dev(b["z_g"]), dev(b["labels"]), apply=False)
    gg = tr.grads("generator")
    state_g = tr.store.export()
    img, dimg = host(ga["fake"])[..., :3], ga["fake"].grad.detach().float().cpu().numpy()[..., :3]
    g_aux = {"fake_logits": host(ga["fake_logits"]), "fake": img}
    ref_shapes = {str(n): tuple(int(s) for s in str(sh).split(",") if s) for n, sh in zip(g["all_names"], g["all_shapes"])}
    assert {k: tuple(v.shape) for k, v in state_g.items()} == ref_shapes
    fp32 = dtype == torch.float32 and not l2
    check_net64(g, float(d_loss.cpu()), gd, d_aux, state_d, float(g_loss.cpu()), gg, g_aux, state_g, tol=tol, gtol=gtol, stol=stol, l2=l2,
                skip=skip, g_gtol=max(gtol, KINK_GTOL) if fp32 else None, g_norm_tol=max(gtol, KINK_NORM_TOL) if fp32 else None)
    if fp32 and split and name in SPLIT_TOL:
        _g_phase_split_check(name, kw, g, state_d, b, img, dimg, gg, *SPLIT_TOL[name])


@pytest.mark.gpu
def test_cuda_network_matches_reference_generator_and_discriminator():
    """One D phase and one G phase of the 64x64 model on the CUDA path (fp32 mode) against the vectors produced by
    BigGAN.generator / BigGAN.discriminator themselves (same variable names and shapes, losses, gradients, state)."""
    _cuda_net("net64")


@pytest.mark.gpu
@UNVERIFIED
@pytest.mark.parametrize("name", ["net128da", "net256", "net512"])
def test_cuda_wiring_of_the_larger_configs_matches_reference(name):
    """BASELINE configs 1-4 as wiring cases (block tables, attention placement, z split) at narrow width, fp32 mode:
    narrow-channel kernels shapes that have not run on hardware yet."""
    _cuda_net(name)


@pytest.mark.gpu
@UNVERIFIED
@pytest.mark.parametrize("name", list(C.NEXT_NETS))
def test_cuda_next_row_networks_match_reference(name):
    """the fork's CLI defaults (prelu, ortho_cosine, flood, DiffAugment) and the layer variants
    (lrelu, resize_conv, conv3, resize_conv1, bn_in_d, ortho) on the CUDA path, fp32 mode."""
    _cuda_net(name)


@pytest.mark.gpu
@UNVERIFIED
@pytest.mark.parametrize("kind", ["ortho", "ortho_cosine"])
def test_cuda_orthogonal_regularizers_match_reference_source(kind):
    """utils.py:180-235: loss and gradient for a conv-shaped and an FC-shaped weight."""
    from helpers import dev_tensor, new_product_store
    from biggan_tensorflow_b200 import utils
    from biggan_tensorflow_b200.tensor import backward
    g = gold("ref_ops.npz")
    for wkey, lkey, gkey, fn in (("w4", "loss4", "g4", utils.orthogonal_regularizer), ("w2", "loss2", "g2", utils.orthogonal_regularizer_fc)):
        new_product_store(torch.float32)
        w = dev_tensor(g[kind + "/" + wkey].astype(np.float32))
        L = fn(1e-4, type=kind)(w)
        backward([(L, torch.ones(1, device="cuda"))])
        torch.cuda.synchronize()
        C.check(g, kind + "/" + lkey, L.numpy().reshape(()), 1e-4)
        C.check(g, kind + "/" + gkey, w.grad.float().cpu().numpy(), 1e-4)


@pytest.mark.gpu
def test_cuda_virtual_batches_and_ema_sampling():
    """utils.virtual_batch_steps (utils.py:242-250): gradients of k virtual batches averaged, one Adam step -- against the
    oracle doing the same; then the inference-mode generator with the EMA weights swapped in against the oracle's.
    (Not gated: only kernels that have run on hardware are involved, and the body passes on the host library.)"""
    _virtual_batches_and_sampling(_to_cuda)


def _virtual_batches_and_sampling(dev):
    from biggan_tensorflow_b200.model import Config
    from biggan_tensorflow_b200.step import Trainer
    from oracle import ref_ops as R
    kw = dict(img_size=32, ch=32, batch_size=4, n_labels=10, sa_gamma_init=0.1)
    cfg, rcfg = Config(virtual_batches=2, **kw), M.Config(**kw)
    bs = [M.synthetic_batch(rcfg, B=4, seed=s) for s in (0, 10)]
    tr = Trainer(cfg, compute_dtype=torch.float32)
    tr.build()
    tr.d_update_virtual([(dev(b["real"]), dev(b["labels"]), dev(b["z_d"]), dev(b["labels"])) for b in bs])
    tr.g_update_virtual([(dev(b["z_g"]), dev(b["labels"])) for b in bs])
    rs = R.RefStore(torch.float64)
    for net, opt_state in (("D", M.AdamState(rcfg.d_lr, rcfg.beta1, rcfg.beta2)),
                           ("G", M.AdamState(rcfg.g_lr, rcfg.beta1, rcfg.beta2, rcfg.moving_decay))):
        acc = None
        for b in bs:
            if net == "D":
                _, gd, _ = M.d_update(rs, rcfg, b["real"], b["labels"], b["z_d"], b["labels"], apply=False)
            else:
                _, gd, _ = M.g_update(rs, rcfg, b["z_g"], b["labels"], apply=False)
            acc = gd if acc is None else {k: acc[k] + gd[k] for k in gd}
        opt_state.apply(rs, {k: v / len(bs) for k, v in acc.items()})
        og = opt_state
    prod = tr.store.export()
    from helpers import assert_close
    for name, v in rs.vars.items():
        if not rs.trainable.get(name, False):
            assert_close(prod[name], v.detach().numpy(), 4e-4, "state " + name)
    # sampling: EMA weights, population statistics
    z, lab = bs[0]["z_g"], bs[0]["labels"]
    img = tr.sample(dev(z), dev(lab), ema=True)
    ema_store = R.RefStore(torch.float64, params={**rs.export(), **{k: v.numpy() for k, v in og.shadow.items()}})
    ref = M.generator(ema_store, rcfg, torch.as_tensor(z, dtype=torch.float64), torch.as_tensor(lab, dtype=torch.float64), is_training=False)
    assert_close(img.numpy()[..., :3], ref.detach().numpy(), 2e-4, "EMA sample")


@pytest.mark.gpu
def test_cuda_checkpoint_round_trip_in_the_reference_format(tmp_path):
    """Trainer -> TensorBundle checkpoint with the swapping saver's names -> fresh Trainer: identical state and next step.
    (Not gated: host code + kernels that have run on hardware; the body passes on the host library.)"""
    _checkpoint_round_trip(tmp_path, _to_cuda)


def _checkpoint_round_trip(tmp_path, dev):
    from biggan_tensorflow_b200 import checkpoint as ck
    from biggan_tensorflow_b200.model import Config
    from biggan_tensorflow_b200.step import Trainer
    kw = dict(img_size=32, ch=32, batch_size=4, n_labels=10, sa_gamma_init=0.1)
    b = M.synthetic_batch(M.Config(**kw), B=4)
    args = (dev(b["real"]), dev(b["labels"]), dev(b["z_d"]), dev(b["z_g"]), dev(b["labels"]))
    tr = Trainer(Config(**kw), compute_dtype=torch.float32)
    tr.train_step(*args)
    tr.train_step(*args)
    prefix = ck.save(tr, str(tmp_path), step=2)
    t = ck.read_bundle(prefix)
    assert "generator/G_logit/kernel/ExponentialMovingAverage" in t and "discriminator/D_logit/kernel/Adam_1" in t
    assert abs(float(t["beta2_power"]) - 0.9 ** 3) < 1e-6 and ck.latest_checkpoint(str(tmp_path)) == prefix
    tr2 = Trainer(Config(**kw), compute_dtype=torch.float32)
    assert ck.load(tr2, prefix) == 2
    a, c = tr.store.export(), tr2.store.export()
    assert sorted(a) == sorted(c) and all(np.array_equal(a[k], c[k]) for k in a)
    assert torch.equal(tr.opt["generator"]["ema"], tr2.opt["generator"]["ema"]) and torch.equal(tr.opt["discriminator"]["v"], tr2.opt["discriminator"]["v"])
    d1, g1 = tr.train_step(*args)
    d2, g2 = tr2.train_step(*args)
    assert abs(float(d1.cpu()) - float(d2.cpu())) <= 1e-6 * abs(float(d1.cpu())) and abs(float(g1.cpu()) - float(g2.cpu())) <= 1e-6 * abs(float(g1.cpu()))


@pytest.mark.gpu
@UNVERIFIED
def test_cuda_input_pipeline_end_to_end(tmp_path):
    """PNG files -> HostBatches (decode on the thread pool, pinned staging) -> DevicePrefetcher (side stream, bg_image_prepare):
    every prepared batch equals the numpy statement of resize + flip + normalise of the decoded files."""
    from test_data_pipeline import _dataset, _tf1_resize_flip_normalise
    from biggan_tensorflow_b200 import data as D
    root, label_file = _dataset(tmp_path, 12, size=10)
    paths, labels = D.load_data("toy", label_file, root=root)
    ref = D.HostBatches(paths, labels, batch_size=4, seed=3, num_threads=2)
    pre = D.DevicePrefetcher(D.HostBatches(paths, labels, batch_size=4, seed=3, num_threads=4), img_size=16, dtype=torch.float32,
                             pad_channels=16, prefetch=3)
    for _ in range(7):
        hb, hl, hf = next(ref)
        img, lab = next(pre)
        assert tuple(img.shape) == (4, 16, 16, 16) and torch.equal(lab.cpu(), torch.from_numpy(hl))
        want = _tf1_resize_flip_normalise(hb, hf, 16, 16)
        assert np.abs(img.cpu().numpy() - want).max() < 2e-6
