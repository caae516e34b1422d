"""TEST INFRASTRUCTURE ONLY.  Generates `tests/golden/ref_*.npz` by executing the reference's OWN source.

    python oracle/make_reference_golden.py            # needs /root/reference (dev container only)

The reference's ops.py, utils.py, DiffAugment_tf.py and BigGAN.py are imported UNMODIFIED from /root/reference
with `oracle/tf1_eager.py` standing in for TensorFlow 1.x (which cannot be installed here).  Every number written
comes out of the reference's functions:

  ref_ops.npz      ops.conv / deconv / fully_connected / cond_bn / bn / self_attention(_2) / resblock_* on the
                   shapes of tests/golden/ref_cases.py: output, gradients w.r.t. inputs and every trainable
                   variable for a seeded output gradient, state after the run (u, pop_mean, ...); spectral_norm
                   alone over two runs; hinge losses + flood (ops.py:753-848); utils.cls_loss_fn('logistic');
                   utils.orthogonal_regularizer(_fc) for 'ortho' and 'ortho_cosine' (utils.py:180-235)
  ref_diffaug.npz  DiffAugment_tf.DiffAugment with the random draws supplied (and logged)
  ref_net*.npz     BigGAN.generator() / BigGAN.discriminator() on a BigGAN object initialised by the reference's own
                   __init__ from main.parse_args() with the canonical flags, at 64 / 128 (+DiffAugment
                   'full') / 256 / 512 pixels with narrow channels, composed into the D and G losses the way
                   build_model() does (BigGAN.py:806-898): losses, logits, fake images, gradient norms of every
                   trainable variable, a few full gradients, every variable's name and shape, all state after each phase

Initial variable values come from the oracle's name-seeded initialiser so that oracle, CUDA path and reference
hold identical weights; inputs from tests/golden/ref_cases.py.  Outputs are stored as float64.
"""
import os
import sys
import types

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = "/root/reference"
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests", "golden"))

from oracle import tf1_eager as E                      # noqa: E402
from oracle.ref_ops import init_array                  # noqa: E402
from oracle import ref_model as M                      # noqa: E402
import ref_cases as C                                  # noqa: E402


def init_fn(full, shape, kind, value, stddev):
    a = init_array(full, shape, kind, value)
    if kind == "trunc_normal" and stddev != 0.02:
        a = (a.astype(np.float64) * (stddev / 0.02)).astype(np.float32)
    return a


def import_reference():
    tf = E.install(init_fn)
    sys.modules.setdefault("imageio", types.ModuleType("imageio"))       # utils.py:4 imports it for PNG output only
    sys.path.insert(0, REF)
    import ops            # noqa: F401  (first: ops.py:16 <-> utils.py:7 import each other)
    import utils          # noqa: F401
    import DiffAugment_tf  # noqa: F401
    import BigGAN         # noqa: F401
    import main           # noqa: F401
    return tf, ops, utils, DiffAugment_tf, BigGAN, main


def fresh(var_init=None):
    E.STATE.reset()
    E.STATE.init_fn = init_fn
    E.STATE.initial_values = {k: np.asarray(v, dtype=np.float32) for k, v in (var_init or {}).items()}


def tens(a, grad=True):
    t = torch.tensor(np.asarray(a), dtype=torch.float64)
    if grad:
        t.requires_grad_(True)
    return E.Tensor(t)


def put(out, key, a):
    """Small arrays are stored whole (float64).  Large ones as a strided sample (float32) plus float64 size / stride /
    sum / L2 norm; tests/golden/ref_cases.py:check() compares both forms."""
    a = np.asarray(a, dtype=np.float64)
    if a.size <= C.WHOLE_LIMIT:
        out[key] = a
        return
    stride = -(-a.size // C.WHOLE_LIMIT)
    flat = a.ravel()
    out[key + "#sub"] = flat[::stride].astype(np.float32)
    out[key + "#stat"] = np.array([a.size, stride, flat.sum(), np.sqrt((flat * flat).sum())], dtype=np.float64)
    out[key + "#shape"] = np.array(a.shape, dtype=np.int64)


def collect(out, prefix, y, inputs, dy):
    """Forward value, gradients for the seeded dy, variables' checksums and post-run state."""
    tvars = [v for v in E.STATE.vars.values() if v.trainable]
    wrt = [t.t for t in inputs.values()] + [v.t for v in tvars]
    grads = torch.autograd.grad((y.t * torch.as_tensor(dy, dtype=torch.float64)).sum(), wrt, allow_unused=True)
    put(out, prefix + "/out", y.t.detach().numpy())
    for (k, _), g in zip(inputs.items(), grads[:len(inputs)]):
        if g is not None:
            put(out, prefix + "/gin/" + k, g.numpy())
    for v, g in zip(tvars, grads[len(inputs):]):
        if g is not None:
            put(out, prefix + "/gvar/" + v.full, g.numpy())
    for v in E.STATE.vars.values():
        out[prefix + "/varsum/" + v.full] = np.array(float(v.t.detach().abs().sum()))
    E.end_run()
    for v in E.STATE.vars.values():
        if not v.trainable:
            out[prefix + "/state/" + v.full] = v.t.detach().numpy().copy()


def gen_cases(ops, cases):
    out = {}
    acts = {"relu": ops.relu, "prelu": ops.prelu, "lrelu": lambda x: ops.lrelu(x, 0.2)}     # BigGAN.py:71-84
    for case in cases:
        fresh(case["var_init"])
        opt = C.make_opt(acts, case["opt"])
        xs = {k: tens(v) for k, v in C.case_inputs(case).items()}
        if case["fn"] in ("prelu", "lrelu"):
            y = getattr(ops, case["fn"])(*xs.values())
        else:
            y = getattr(ops, case["fn"])(*xs.values(), **case["kw"], opt=opt)
        if case["relu_after"]:
            y = opt["act"](y)
        collect(out, case["name"], y, xs, C.case_dy(case, tuple(y.shape)))
    return out


def gen_ops(ops, utils):
    out = gen_cases(ops, C.OP_CASES)
    tf = sys.modules["tensorflow"]
    for K, Cc in C.SN_CASES:                                      # spectral_norm alone, two consecutive runs
        fresh()
        name = "sn_%dx%d" % (K, Cc)
        for run in (1, 2):
            with tf.variable_scope("t", reuse=(run == 2)):
                w = tf.get_variable("kernel", [K, Cc], initializer=ops.weight_init)
                wn = ops.spectral_norm(w)
            dy = np.random.default_rng(1000 + run).standard_normal((K, Cc)).astype(np.float32)
            (g,) = torch.autograd.grad((wn.t * torch.as_tensor(dy, dtype=torch.float64)).sum(), [w.t])
            put(out, "%s/run%d/out" % (name, run), wn.t.detach().numpy())
            put(out, "%s/run%d/gvar" % (name, run), g.numpy())
            E.end_run()
            out["%s/run%d/u" % (name, run)] = E.STATE.vars["t/u"].t.detach().numpy().copy()
    rng = np.random.default_rng(14)                              # hinge losses (+ flood)
    for flood in C.LOSS_FLOODS:
        real = rng.standard_normal((16, 1)).astype(np.float32) * 2
        fake = rng.standard_normal((16, 1)).astype(np.float32) * 2
        tr, tfk = tens(real), tens(fake)
        L = ops.discriminator_loss('hinge', real=tr, fake=tfk, flood_level=flood)
        gr, gf = torch.autograd.grad(L.t, [tr.t, tfk.t])
        tf2 = tens(fake)
        G = ops.generator_loss('hinge', fake=tf2, real=tr, flood_level=flood)
        (gg,) = torch.autograd.grad(G.t, [tf2.t])
        p = "hinge_flood_%g" % flood
        out.update({p + "/real": real, p + "/fake": fake, p + "/d_loss": L.t.detach().numpy(), p + "/d_greal": gr.numpy(),
                    p + "/d_gfake": gf.numpy(), p + "/g_loss": G.t.detach().numpy(), p + "/g_gfake": gg.numpy()})
    rng = np.random.default_rng(15)                              # class loss, utils.py:366-369
    logits = (rng.standard_normal((6, 10)) * 3).astype(np.float32)
    labels = np.eye(10, dtype=np.float32)[rng.integers(0, 10, 6)]
    wts = rng.random(10).astype(np.float32) + 0.5
    tl = tens(logits)
    L = utils.cls_loss_fn('logistic', tens(wts, False))(tens(labels, False), tl)
    (g,) = torch.autograd.grad(L.t, [tl.t])
    out.update({"cls_logistic/logits": logits, "cls_logistic/labels": labels, "cls_logistic/weights": wts,
                "cls_logistic/loss": L.t.detach().numpy(), "cls_logistic/glogits": g.numpy()})
    rng = np.random.default_rng(17)                              # orthogonal regularisers, utils.py:180-235
    for kind in ("ortho", "ortho_cosine"):
        w4 = (rng.standard_normal((3, 3, 8, 12)) * 0.1).astype(np.float32)
        w2 = (rng.standard_normal((40, 12)) * 0.1).astype(np.float32)
        t4, t2 = tens(w4), tens(w2)
        l4 = utils.orthogonal_regularizer(1e-4, type=kind)(t4)
        l2 = utils.orthogonal_regularizer_fc(1e-4, type=kind)(t2)
        g4, = torch.autograd.grad(l4.t, [t4.t])
        g2, = torch.autograd.grad(l2.t, [t2.t])
        out.update({kind + "/w4": w4, kind + "/w2": w2, kind + "/loss4": l4.t.detach().numpy(), kind + "/loss2": l2.t.detach().numpy(),
                    kind + "/g4": g4.numpy(), kind + "/g2": g2.numpy()})
    return out


def gen_diffaug(DA):
    out = {}
    for name, policy, shape in C.DIFFAUG_CASES:
        fresh()
        B, H, W, _ = shape
        rng = np.random.Generator(np.random.PCG64(4))
        from oracle.ref_ops import draw_diffaug_randoms
        rnd = draw_diffaug_randoms(rng, B, H, W)
        order = {"color": ["brightness", "saturation", "contrast"], "translation": ["translation_x", "translation_y"],
                 "cutout": ["cutout_x", "cutout_y"]}
        E.STATE.random_queue = [rnd[k] for p in policy.split(",") for k in order[p]]
        x, dy = C.diffaug_inputs(shape)
        tx = tens(x)
        y = DA.DiffAugment(tx, policy=policy)
        assert not E.STATE.random_queue, "the reference drew fewer random tensors than supplied"
        (g,) = torch.autograd.grad((y.t * torch.as_tensor(dy, dtype=torch.float64)).sum(), [tx.t])
        put(out, name + "/out", y.t.detach().numpy())
        put(out, name + "/gx", g.numpy())
        for k, v in rnd.items():
            out[name + "/rnd/" + k] = v
        # the ranges the reference asked tf.random.uniform for, in call order (pins draw_diffaug_randoms)
        out[name + "/ranges"] = np.array([[lo, hi] for (_, lo, hi, _, _) in E.STATE.random_log], dtype=np.float64)
    return out


def gen_net(kw, BigGAN, main, utils, ops, DA):
    da = kw.get("da_policy", "")
    argv = ["main.py", "--img_size", str(kw["img_size"]), "--ch", str(kw["ch"]), "--batch_size", str(kw["batch_size"]),
            "--n_labels", str(kw["n_labels"]), "--z_dim", str(kw["z_dim"]),
            # canonical flag set, SURVEY.md 8(d)
            "--gan_type", "hinge", "--activation", kw.get("activation", "relu"), "--sn", "true", "--bn_type", "batch_norm",
            "--bn_momentum", "0.98", "--conv_padding", "reflect", "--upsampling_method", kw.get("upsampling_method", "deconv4"),
            "--g_conv", kw.get("g_conv", "deconv3"), "--downsampling_method", kw.get("downsampling_method", "strided_conv3"),
            "--bias_in_sa", "true", "--bias_in_d", "false", "--bn_in_d", "true" if kw.get("bn_in_d") else "false",
            "--first_split_ratio", "3", "--g_first_level_dense_layer", "true",
            "--g_regularization", kw.get("g_regularization", "none"), "--g_regularization_factor", str(kw.get("g_regularization_factor", 1e-4)),
            "--d_flood", str(kw.get("d_flood", 0)), "--g_flood", str(kw.get("g_flood", 0)), "--da_policy", "full" if da else "none"]
    old = sys.argv
    sys.argv = argv
    try:
        args = main.parse_args()
    finally:
        sys.argv = old
    if not da:
        args.da_policy = ""
    # BigGAN.__init__ (BigGAN.py:14-231) reads the dataset directory and creates the sample folder: neutralise the two
    # filesystem helpers it took from utils, run the reference's own flag processing
    BigGAN.load_data = lambda **k: ([], [])
    BigGAN.check_folder = lambda d: d
    gan = BigGAN.BigGAN(None, args)
    assert gan.da_policy == da, (gan.da_policy, da)              # GANBase.py:53-55 expands 'full'
    cfg = M.Config(**kw)
    batch = M.synthetic_batch(cfg, B=kw["batch_size"])
    fresh({"generator/self_attention/gamma": [kw["sa_gamma_init"]], "discriminator/self_attention/gamma": [kw["sa_gamma_init"]]})
    out = {}
    if da:                                                       # draws for D(real), D(fake) [D phase], D(fake) [G phase]
        from oracle.ref_ops import draw_diffaug_randoms
        rng = np.random.Generator(np.random.PCG64(5))
        order = ["brightness", "saturation", "contrast", "translation_x", "translation_y", "cutout_x", "cutout_y"]
        for i in range(3):
            rnd = draw_diffaug_randoms(rng, kw["batch_size"], kw["img_size"], kw["img_size"])
            E.STATE.random_queue += [rnd[k] for k in order]
            for k in order:
                out["rnd%d/%s" % (i, k)] = rnd[k]
    cls_loss = utils.cls_loss_fn(gan.cls_loss_type, tens(gan.cls_loss_weights, False))

    def names(sub):
        return [v.full for v in E.STATE.vars.values() if v.trainable and sub in v.full]

    def pick(nl):
        """Full gradients kept: the first and last three trainables, the attention gamma and its f_conv kernel."""
        return sorted(set(nl[:3] + nl[-3:] + [n for n in nl if n.endswith("self_attention/gamma") or n.endswith("f_conv/kernel")]))

    def state(tag):
        for v in E.STATE.vars.values():
            if not v.trainable:
                put(out, tag + v.full, v.t.detach().numpy())

    # ---- D phase: BigGAN.py:806-807, :853, :857-883 -------------------------------------------------
    real = tens(batch["real"], False)
    labels = tens(batch["labels"], False)
    d_real = gan.discriminator(DA.DiffAugment(real, policy=gan.da_policy))
    fake = gan.generator(tens(batch["z_d"], False), labels)
    d_fake = gan.discriminator(DA.DiffAugment(fake, policy=gan.da_policy), reuse=True)
    d_cls = gan.d_cls_loss_weight * cls_loss(labels, d_real["cls"])
    d_loss = ops.discriminator_loss(gan.d_loss_func, real=d_real["real"], fake=d_fake["real"], flood_level=gan.d_flood) + d_cls
    dn = names("discriminator")
    grads = torch.autograd.grad(d_loss.t, [E.STATE.vars[n].t for n in dn], allow_unused=True)
    out["D_names"] = np.array(dn)
    out["G_names"] = np.array(names("generator"))
    out["all_names"] = np.array(list(E.STATE.vars.keys()))
    out["all_shapes"] = np.array([",".join(str(s) for s in v.shape) for v in E.STATE.vars.values()])
    out["d_loss"] = d_loss.t.detach().numpy()
    out["d_cls_loss"] = d_cls.t.detach().numpy()
    out["d_real_logits"] = d_real["real"].t.detach().numpy()
    out["d_fake_logits"] = d_fake["real"].t.detach().numpy()
    put(out, "d_fake_images", fake.t.detach().numpy())
    out["D_grad_norms"] = np.array([0.0 if g is None else float(g.norm()) for g in grads])
    gd = dict(zip(dn, grads))
    for n in pick(dn):
        put(out, "D_grad/" + n, gd[n].numpy())
    E.end_run()
    state("state_after_D/")

    # ---- G phase (no optimizer step in between): BigGAN.py:883, :857, :894-896 ------------------------
    fake = gan.generator(tens(batch["z_g"], False), labels, reuse=True)
    d_fake = gan.discriminator(DA.DiffAugment(fake, policy=gan.da_policy), reuse=True)
    g_cls = gan.g_cls_loss_weight * cls_loss(labels, d_fake["cls"])
    g_loss = ops.generator_loss(gan.gan_type, fake=d_fake["real"], real=d_real["real"], flood_level=gan.g_flood) + g_cls
    if gan.g_regularization_method != 'none':                    # BigGAN.py:897-898
        tf = sys.modules["tensorflow"]
        regs = tf.losses.get_regularization_losses()
        out["n_regularization_losses"] = np.array(len(regs))
        out["regularization_loss_sum"] = tf.add_n(regs).t.detach().numpy()
        g_loss = tf.add_n([g_loss] + regs)
    gn = names("generator")
    grads = torch.autograd.grad(g_loss.t, [E.STATE.vars[n].t for n in gn], allow_unused=True)
    out["g_loss"] = g_loss.t.detach().numpy()
    out["g_cls_loss"] = g_cls.t.detach().numpy()
    out["g_fake_logits"] = d_fake["real"].t.detach().numpy()
    put(out, "g_fake_images", fake.t.detach().numpy())
    out["G_grad_norms"] = np.array([0.0 if g is None else float(g.norm()) for g in grads])
    gg = dict(zip(gn, grads))
    for n in pick(gn):
        put(out, "G_grad/" + n, gg[n].numpy())
    E.end_run()
    assert not E.STATE.random_queue
    state("state_after_G/")
    return out


def main_():
    tf, ops, utils, DA, BigGAN, main = import_reference()
    gold = os.path.join(ROOT, "tests", "golden")
    only = sys.argv[1:]
    jobs = [("ref_ops.npz", lambda: gen_ops(ops, utils)), ("ref_diffaug.npz", lambda: gen_diffaug(DA))]
    jobs += [("ref_ops_next.npz", lambda: gen_cases(ops, C.NEXT_OP_CASES))]
    jobs += [("ref_%s.npz" % name, (lambda kw=kw: gen_net(kw, BigGAN, main, utils, ops, DA)))
             for name, kw in list(C.NETS.items()) + list(C.NEXT_NETS.items())]
    for fname, job in jobs:
        if only and fname not in only:
            continue
        data = job()
        np.savez_compressed(os.path.join(gold, fname), **data)
        print(fname, len(data), "arrays", os.path.getsize(os.path.join(gold, fname)) // 1024, "KiB")


if __name__ == "__main__":
    main_()
