"""Golden vectors produced by the REFERENCE'S OWN SOURCE (`tests/golden/ref_*.npz`, written by
`oracle/make_reference_golden.py`: ops.py / utils.py / DiffAugment_tf.py / BigGAN.py imported unmodified from
/root/reference and executed over the eager TF-1 stand-in `oracle/tf1_eager.py`).

This file (CPU, `-m "not gpu"`): the oracle restatement reproduces them in float64 -- outputs, gradients, variable names
and shapes, state after the run.  `tests/test_zz_reference_golden_gpu.py` (`-m gpu`): the CUDA path (through the C ABI)
reproduces them within north_star's tolerances (1e-4 relative in fp32 mode, 2e-2 in bf16 mode), DiffAugment index work
bit-exactly.
"""
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
from oracle import ref_diffaug_np as RNP    # noqa: E402

F64_TOL = 2e-9          # oracle vs reference source, both float64 (large arrays: sampled in float32, +1e-7)


def gold(name):
    return np.load(os.path.join(HERE, "golden", name))


# ------------------------------------------------------------------------------------------------
# oracle side (CPU)
# ------------------------------------------------------------------------------------------------

def oracle_run_case(case):
    rs = R.RefStore(torch.float64, params={k: np.asarray(v, np.float32) for k, v in case["var_init"].items()})
    acts = {"relu": R.relu, "prelu": R.make_prelu(rs), "lrelu": lambda x: R.lrelu(x, 0.2)}
    opt = C.make_opt(acts, case["opt"])
    xs = {k: torch.tensor(v, dtype=torch.float64, requires_grad=True) for k, v in C.case_inputs(case).items()}
    if case["fn"] in ("prelu", "lrelu"):
        y = acts[case["fn"]](*xs.values())
    else:
        y = getattr(R, case["fn"])(rs, *xs.values(), **case["kw"], opt=opt)
    if case["relu_after"]:
        y = opt["act"](y)
    return rs, xs, y


ALL_OP_CASES = [(c, "ref_ops.npz") for c in C.OP_CASES] + [(c, "ref_ops_next.npz") for c in C.NEXT_OP_CASES]


@pytest.mark.parametrize("case,fname", ALL_OP_CASES, ids=[c["name"] for c, _ in ALL_OP_CASES])
def test_oracle_op_matches_reference_source(case, fname):
    g = gold(fname)
    p = case["name"]
    rs, xs, y = oracle_run_case(case)
    # identical variables: names, and the checksum of every initial value (state variables were replaced by the call)
    ref_names = sorted(k.split("/varsum/")[1] for k in g.files if k.startswith(p + "/varsum/"))
    assert sorted(rs.vars.keys()) == ref_names
    tnames = [n for n in rs.vars if rs.trainable.get(n, False)]
    for n in tnames:
        assert abs(float(rs.vars[n].detach().abs().sum()) - float(g[p + "/varsum/" + n])) <= 1e-9 * max(1.0, float(g[p + "/varsum/" + n])), n
    dy = C.case_dy(case, tuple(y.shape))
    wrt = list(xs.values()) + [rs.vars[n] for n in tnames]
    grads = torch.autograd.grad((y * torch.as_tensor(dy, dtype=torch.float64)).sum(), wrt, allow_unused=True)
    C.check(g, p + "/out", y.detach().numpy(), F64_TOL)
    for k, gr in zip(xs.keys(), grads[:len(xs)]):
        assert (gr is not None) == ((p + "/gin/" + k in g.files) or (p + "/gin/" + k + "#stat" in g.files)), k
        if gr is not None:
            C.check(g, p + "/gin/" + k, gr.numpy(), F64_TOL)
    n_checked = 0
    gkeys = sorted(set(k.split("#")[0] for k in g.files if k.startswith(p + "/gvar/")))
    gscale = max([C.golden_absmax(g, k) for k in gkeys] + [0.0])
    for n, gr in zip(tnames, grads[len(xs):]):
        key = p + "/gvar/" + n
        has = key in g.files or key + "#stat" in g.files
        assert (gr is not None) == has, n
        if gr is not None:
            if C.golden_absmax(g, key) <= 1e-10 * gscale:
                # analytically zero on both sides, only rounding noise: d/d f_conv bias (softmax shift invariance),
                # d/d bias of a conv that feeds a batch norm (the mean is removed)
                assert float(gr.abs().max()) <= 1e-9 * gscale, n
                continue
            C.check(g, key, gr.numpy(), F64_TOL)
            n_checked += 1
    assert n_checked > 0 or not tnames
    for n, v in rs.vars.items():
        if not rs.trainable.get(n, False):
            C.check(g, p + "/state/" + n, v.detach().numpy(), F64_TOL)


@pytest.mark.parametrize("K,Cc", C.SN_CASES)
def test_oracle_spectral_norm_two_runs_match_reference_source(K, Cc):
    """ops.py:718-747 evaluated in two consecutive runs: W/sigma, its gradient (u_hat, v_hat frozen), u after each run."""
    g = gold("ref_ops.npz")
    rs = R.RefStore(torch.float64)
    for run in (1, 2):
        rs.begin_phase()
        with rs.variable_scope("t"):
            w = rs.get_variable("kernel", [K, Cc])
            wn = R.spectral_norm(rs, w)
        dy = np.random.default_rng(1000 + run).standard_normal((K, Cc)).astype(np.float32)
        (gw,) = torch.autograd.grad((wn * torch.as_tensor(dy, dtype=torch.float64)).sum(), [w])
        p = "sn_%dx%d/run%d" % (K, Cc, run)
        C.check(g, p + "/out", wn.detach().numpy(), F64_TOL)
        C.check(g, p + "/gvar", gw.numpy(), F64_TOL)
        C.check(g, p + "/u", rs.vars["t/u"].numpy(), F64_TOL)


def test_oracle_losses_match_reference_source():
    g = gold("ref_ops.npz")
    for flood in C.LOSS_FLOODS:
        p = "hinge_flood_%g" % flood
        rr = torch.tensor(g[p + "/real"], dtype=torch.float64, requires_grad=True)
        rf = torch.tensor(g[p + "/fake"], dtype=torch.float64, requires_grad=True)
        L = R.discriminator_loss("hinge", rr, rf, flood)
        gr, gf = torch.autograd.grad(L, [rr, rf])
        G = R.generator_loss("hinge", rf, None, flood)
        (gg,) = torch.autograd.grad(G, [rf])
        for key, v in (("d_loss", L), ("d_greal", gr), ("d_gfake", gf), ("g_loss", G), ("g_gfake", gg)):
            C.check(g, p + "/" + key, v.detach().numpy(), 1e-12)
    rl = torch.tensor(g["cls_logistic/logits"], dtype=torch.float64, requires_grad=True)
    L = R.cls_loss_logistic(torch.tensor(g["cls_logistic/labels"], dtype=torch.float64), rl,
                            torch.tensor(g["cls_logistic/weights"], dtype=torch.float64))
    (gl,) = torch.autograd.grad(L, [rl])
    C.check(g, "cls_logistic/loss", L.detach().numpy(), 1e-12)
    C.check(g, "cls_logistic/glogits", gl.numpy(), 1e-12)


@pytest.mark.parametrize("kind", ["ortho", "ortho_cosine"])
def test_oracle_orthogonal_regularizers_match_reference_source(kind):
    """utils.py:180-235."""
    g = gold("ref_ops.npz")
    for wkey, lkey, gkey, fn in (("w4", "loss4", "g4", R.orthogonal_regularizer), ("w2", "loss2", "g2", R.orthogonal_regularizer_fc)):
        w = torch.tensor(g[kind + "/" + wkey], dtype=torch.float64, requires_grad=True)
        L = fn(1e-4, type=kind)(w)
        (gw,) = torch.autograd.grad(L, [w])
        C.check(g, kind + "/" + lkey, L.detach().numpy(), 1e-11)
        C.check(g, kind + "/" + gkey, gw.numpy(), 1e-11)


@pytest.mark.parametrize("name,policy,shape", C.DIFFAUG_CASES, ids=[c[0] for c in C.DIFFAUG_CASES])
def test_oracle_diffaugment_matches_reference_source(name, policy, shape):
    """DiffAugment_tf.py:8-73 with the same random draws: values, gradient, and the ranges the reference asks
    tf.random.uniform for (which is what `draw_diffaug_randoms` must reproduce)."""
    g = gold("ref_diffaug.npz")
    x, dy = C.diffaug_inputs(shape)
    rnd = {k.split("/rnd/")[1]: g[k] for k in g.files if k.startswith(name + "/rnd/")}
    tx = torch.tensor(x, dtype=torch.float64, requires_grad=True)
    y = R.DiffAugment(tx, policy=policy, rnd=R.rnd_to_torch(rnd))
    (gx,) = torch.autograd.grad((y * torch.as_tensor(dy, dtype=torch.float64)).sum(), [tx])
    C.check(g, name + "/out", y.detach().numpy(), 1e-12)
    C.check(g, name + "/gx", gx.numpy(), 1e-12)
    B, H, W, _ = shape
    sh, cs = R.translation_shift(H), R.cutout_size(H)
    want = {"color": [[0, 1]] * 3, "translation": [[-sh, sh + 1]] * 2, "cutout": [[0, H + (1 - cs % 2)]] * 2}
    assert g[name + "/ranges"].tolist() == [list(map(float, r)) for p in policy.split(",") for r in want[p]]
    if policy in ("translation", "cutout"):           # the literal numpy restatement of the index ops agrees bit for bit
        f = RNP.rand_translation_np if policy == "translation" else RNP.rand_cutout_np
        a, b = ("translation_x", "translation_y") if policy == "translation" else ("cutout_x", "cutout_y")
        ynp = f(x.astype(np.float64), rnd[a], rnd[b])
        C.check(g, name + "/out", ynp, 0.0)


def _oracle_net(name):
    kw = C.NETS[name] if name in C.NETS else C.NEXT_NETS[name]
    g = gold("ref_%s.npz" % name)
    cfg = M.Config(**kw)
    b = M.synthetic_batch(cfg, B=kw["batch_size"])
    rs = R.RefStore(torch.float64)
    rnd = [None] * 3
    if kw.get("da_policy"):
        rnd = [R.rnd_to_torch({k.split("/")[1]: g[k] for k in g.files if k.startswith("rnd%d/" % i)}) for i in range(3)]
    d_loss, gd, d_aux = M.d_update(rs, cfg, b["real"], b["labels"], b["z_d"], b["labels"], rnd_real=rnd[0], rnd_fake=rnd[1], apply=False)
    state_d = {k: v.detach().numpy().copy() for k, v in rs.vars.items()}
    g_loss, gg, g_aux = M.g_update(rs, cfg, b["z_g"], b["labels"], rnd_fake=rnd[2], apply=False)
    state_g = {k: v.detach().numpy().copy() for k, v in rs.vars.items()}
    return g, rs, d_loss, gd, d_aux, state_d, g_loss, gg, g_aux, state_g


def check_net64(g, d_loss, gd, d_aux, state_d, g_loss, gg, g_aux, state_g, tol, gtol, stol, l2=False, skip=(), g_gtol=None,
                g_norm_tol=None):
    """Shared by the CPU (oracle) and GPU (CUDA path) tests.  gd / gg: name -> gradient array.  g_gtol / g_norm_tol:
    tolerance of the G-PHASE gradients (elements / norms) when it differs from gtol (fp32 CUDA path: the end-to-end
    comparison through D(G(z)) is limited by ReLU / max-pool ties, see tests/test_zz_reference_golden_gpu.py)."""
    C.check(g, "d_loss", np.reshape(d_loss, ()), tol)
    C.check(g, "g_loss", np.reshape(g_loss, ()), tol)
    if d_aux is not None:
        C.check(g, "d_real_logits", d_aux["real_logits"], tol, l2=l2)
        C.check(g, "d_fake_logits", d_aux["fake_logits"], tol, l2=l2)
        C.check(g, "d_fake_images", d_aux["fake"], tol, l2=l2)
        C.check(g, "g_fake_logits", g_aux["fake_logits"], tol, l2=l2)
        C.check(g, "g_fake_images", g_aux["fake"], tol, l2=l2)
    for which, grads in (("D", gd), ("G", gg)):
        names = [str(n) for n in g[which + "_names"]]
        assert sorted(names) == sorted(grads.keys()), sorted(set(names) ^ set(grads.keys()))
        ref = g[which + "_grad_norms"]
        norms = np.array([float(np.linalg.norm(np.asarray(grads[n], dtype=np.float64))) for n in names])
        ok = ref > 1e-3 * ref.max()                      # analytically-zero gradients (f_conv bias) hold only noise
        etol = g_gtol if (which == "G" and g_gtol is not None) else gtol
        ntol = g_norm_tol if (which == "G" and g_norm_tol is not None) else gtol
        np.testing.assert_allclose(norms[ok], ref[ok], rtol=ntol, err_msg=which + " gradient norms")
        for k in sorted(set(f.split("#")[0] for f in g.files if f.startswith(which + "_grad/"))):
            n = k[len(which) + 6:]
            if n.endswith("f_conv/bias") or n.endswith(tuple(skip)):
                continue
            C.check(g, k, grads[n], etol, l2=l2)
    for tag, st in (("state_after_D/", state_d), ("state_after_G/", state_g)):
        for k in sorted(set(f.split("#")[0] for f in g.files if f.startswith(tag))):
            C.check(g, k, st[k[len(tag):]], stol, l2=l2)


@pytest.mark.parametrize("name", list(C.NETS) + list(C.NEXT_NETS))
def test_oracle_network_matches_reference_generator_and_discriminator(name):
    """BigGAN.generator / BigGAN.discriminator (canonical flags; 64, 128 + DiffAugment, 256, 512 pixels) composed as build_model
    does: the oracle creates the same variables (names, shapes) and reproduces losses, logits, images, gradients, state."""
    g, rs, d_loss, gd, d_aux, state_d, g_loss, gg, g_aux, state_g = _oracle_net(name)
    ref_shapes = {str(n): tuple(int(s) for s in str(sh).split(",") if s) for n, sh in zip(g["all_names"], g["all_shapes"])}
    assert {k: tuple(v.shape) for k, v in rs.vars.items()} == ref_shapes
    check_net64(g, d_loss, {k: v.numpy() for k, v in gd.items()}, {k: v.numpy() for k, v in d_aux.items()}, state_d,
                g_loss, {k: v.numpy() for k, v in gg.items()}, {k: v.numpy() for k, v in g_aux.items()}, state_g,
                tol=F64_TOL, gtol=1e-7, stol=F64_TOL)
    if "regularization_loss_sum" in g.files:              # BigGAN.py:897-898: one regularisation loss per generator weight
        regs = rs.regularization_losses()
        assert len(regs) == int(g["n_regularization_losses"])
        C.check(g, "regularization_loss_sum", float(sum(regs).detach()), 1e-10)
