"""CPU ORACLE -- TEST INFRASTRUCTURE ONLY (see oracle/ref_ops.py header: parity pinned to the reference's source
through tests/golden/ref_net64.npz, TensorFlow's kernel layer restated).

Restates the default-flag wiring of the reference's generator / discriminator / loss / optimizer
(BigGAN.py:246-582, :591-715, :790-930, :1061-1084; utils.py:252-267) on top of oracle/ref_ops.py.
Canonical flag set: SURVEY.md section 8(d).
"""
import copy
from dataclasses import dataclass, field

import numpy as np
import torch

from . import ref_ops as R


@dataclass
class Config:
    img_size: int = 128
    ch: int = 96
    batch_size: int = 256
    n_labels: int = 1000
    z_dim: int = 256
    c_dim: int = 3
    first_split_ratio: int = 3
    da_policy: str = ""
    gan_type: str = "hinge"
    d_flood: float = 0.0
    g_flood: float = 0.0
    g_lr: float = 5e-5
    d_lr: float = 2e-4
    beta1: float = 0.0
    beta2: float = 0.9
    moving_decay: float = 0.999
    bn_momentum: float = 0.98
    d_cls_loss_weight: float = 5.0
    g_cls_loss_weight: float = 1.0
    sa_gamma_init: float = 0.0     # the reference initialises to 0 (ops.py:486); parity tests use 0.1
    # SURVEY 8(f) "next" rows: the fork's CLI-default extras and layer variants (main.py:49-67)
    activation: str = "relu"                   # relu | prelu | lrelu            (BigGAN.py:71-84)
    upsampling_method: str = "deconv4"         # ops.py:200-218
    downsampling_method: str = "strided_conv3"  # ops.py:269-291
    g_conv: str = "deconv3"                    # ops.py:220-230
    bn_in_d: bool = False
    g_regularization: str = "none"             # none | ortho | ortho_cosine    (BigGAN.py:257-271)
    g_regularization_factor: float = 1e-4

    @property
    def depth(self):
        return self.img_size.bit_length() - 2                                   # BigGAN.py:19

    @property
    def acgan(self):
        return self.n_labels > 0


CONFIGS = {
    0: dict(img_size=32, ch=32, batch_size=64, n_labels=10),
    1: dict(img_size=128, ch=96, batch_size=256, n_labels=1000),
    2: dict(img_size=128, ch=96, batch_size=128, n_labels=1000, da_policy="color,translation,cutout"),
    3: dict(img_size=256, ch=96, batch_size=256, n_labels=1000),
    4: dict(img_size=512, ch=64, batch_size=64, n_labels=1000),
}


def round_up(val, multiple):
    return (int(val) + multiple - 1) // multiple * multiple                     # BigGAN.py:236-237


def g_block_info(img_size):
    """BigGAN.py:290-294 (+ the 32x32 extrapolation of SURVEY A.1)."""
    if img_size == 32:
        return {"counts": [1, 1, 1], "sa_index": 2}
    if img_size == 64:
        return {"counts": [1, 1, 1, 1], "sa_index": 3}
    if img_size == 128:
        return {"counts": [1, 1, 1, 1, 1], "sa_index": 4}
    if img_size == 256:
        return {"counts": [1, 2, 1, 1, 1], "sa_index": 3}
    if img_size == 512:
        return {"counts": [1, 2, 1, 1, 2], "sa_index": 3}
    raise ValueError("Invalid image size specified: " + str(img_size))


def d_block_info(img_size):
    """BigGAN.py:607-611 (+ 32x32 extrapolation)."""
    if img_size == 32:
        return {"counts": [1, 1, 1], "sa_index": 1}
    if img_size == 64:
        return {"counts": [1, 1, 1, 1], "sa_index": 1}
    if img_size == 128:
        return {"counts": [1, 1, 1, 1, 1], "sa_index": 1}
    if img_size == 256:
        return {"counts": [1, 1, 1, 2, 1], "sa_index": 2}
    if img_size == 512:
        return {"counts": [1, 2, 1, 1, 2], "sa_index": 2}
    raise ValueError("Invalid image size specified: " + str(img_size))


def activation_fn(store, cfg):
    """BigGAN.py:71-84."""
    if cfg.activation == "relu":
        return R.relu
    if cfg.activation == "prelu":
        return R.make_prelu(store)
    if cfg.activation == "lrelu":
        return lambda x: R.lrelu(x, 0.2)
    raise ValueError("Unknown activation function: " + str(cfg.activation))


def make_opts(cfg, is_training=True, store=None):
    bn_options = {"type": "batch_norm", "momentum": cfg.bn_momentum}            # GANBase.py:39-41
    conv_options = {"padding_type": "reflect", "sn": True}                      # GANBase.py:49-51
    act = activation_fn(store, cfg)
    g_opt = {"sn": True, "is_training": is_training, "upsampling_method": cfg.upsampling_method,
             "g_conv": cfg.g_conv, "act": act, "self_attention_bias": True,
             "bn": copy.deepcopy(bn_options), "conv": copy.deepcopy(conv_options),
             "fc_regularizer": None}                                            # BigGAN.py:248-274
    g_opt["conv"]["regularizer"] = None
    if is_training and cfg.g_regularization != "none":                          # BigGAN.py:257-271
        if cfg.g_regularization not in ("ortho", "ortho_cosine"):
            raise ValueError("Unknown regularization method: " + str(cfg.g_regularization))
        g_opt["conv"]["regularizer"] = R.orthogonal_regularizer(cfg.g_regularization_factor, type=cfg.g_regularization)
        g_opt["fc_regularizer"] = R.orthogonal_regularizer_fc(cfg.g_regularization_factor, type=cfg.g_regularization)
    d_opt = {"sn": True, "is_training": is_training, "bn_in_d": cfg.bn_in_d, "act": act,
             "downsampling_method": cfg.downsampling_method, "self_attention_bias": True,
             "bn": copy.deepcopy(bn_options), "conv": copy.deepcopy(conv_options)}  # BigGAN.py:593-600
    return g_opt, d_opt


def z_split_sizes(cfg):
    """BigGAN.py:281-288."""
    if cfg.first_split_ratio > 1:
        split_dim = cfg.z_dim // (cfg.depth - 1 + cfg.first_split_ratio)
        first_split_dim = cfg.z_dim - (cfg.depth - 1) * split_dim
    else:
        split_dim = cfg.z_dim // cfg.depth
        first_split_dim = split_dim
    return [first_split_dim] + [split_dim] * (cfg.depth - 1)


def generator(store, cfg, z, cls_z, is_training=True):
    """BigGAN.py:246-582, canonical flags.  z [B,1,1,z_dim], cls_z [B,n_labels]."""
    opt, _ = make_opts(cfg, is_training, store)
    with store.variable_scope("generator"):
        split_sizes = z_split_sizes(cfg)
        block_info = g_block_info(cfg.img_size)
        z_split = list(torch.split(z, split_sizes, dim=-1))                     # :335
        if cfg.acgan:
            cz = cls_z.reshape(-1, 1, 1, cfg.n_labels)                          # :358
            z_split = [torch.cat([s, cz], dim=-1) for s in z_split]             # :361-365
        n_blocks = len(block_info["counts"])
        ch_mul = 2 ** (n_blocks - 1)                                            # :427
        ch = round_up(int(cfg.ch * 2.0 ** (n_blocks - 0 - 1)), 8)               # :428, :242-243
        zi = 0
        f_width = round_up((split_sizes[0] + cfg.n_labels) * 1.85, 8)           # :433
        if opt["act"] == R.relu:                                                # :434 (no scope, backward compatibility)
            x = R.fully_connected(store, z_split[0], units=f_width, scope="dense1", opt=opt)   # :436
            x = R.relu(x)
            x = R.fully_connected(store, x, units=4 * 4 * ch, scope="dense2", opt=opt)         # :438
        else:
            with store.variable_scope("first"):                                 # :440-443
                x = R.fully_connected(store, z_split[0], units=f_width, scope="dense1", opt=opt)
                x = opt["act"](x)
                x = R.fully_connected(store, x, units=4 * 4 * ch, scope="dense2", opt=opt)
        zi += 1
        x = x.reshape(-1, 4, 4, ch)                                             # :446
        b_i = 0
        for block_count in block_info["counts"]:
            scope = "resblock_up_" + str(ch_mul)                                # :450
            for sb_i in range(block_count):
                layer_z = z_split[zi]
                zi += 1
                if block_count > 1:
                    scope = scope + "_" + str(sb_i)                             # :455
                x = R.resblock_up_condition(store, x, layer_z, channels=ch, use_bias=False,
                                            opt=opt, scope=scope)               # :479
            b_i += 1
            if b_i == block_info["sa_index"]:
                x = _sa2(store, cfg, x, ch, opt)                                # :482-483
            ch = round_up(int(cfg.ch * 2.0 ** (n_blocks - b_i - 1)), 8)         # :488
            ch_mul = ch_mul // 2
        x = R.bn(store, x, opt=opt)                                             # :491
        x = opt["act"](x)                                                       # :492
        x = R.conv(store, x, channels=cfg.c_dim, kernel=3, stride=1, pad=1, use_bias=False,
                   opt=opt, scope="G_logit")                                    # :570
        return R.tanh(x)                                                        # :580


def _sa2(store, cfg, x, ch, opt):
    # attention gamma is created by self_attention_2 itself; pre-create it with the config's init
    with store.variable_scope("self_attention"):
        store.get_variable("gamma", [1], kind="const", value=cfg.sa_gamma_init)
    return R.self_attention_2(store, x, channels=ch, opt=opt, scope="self_attention")


def discriminator(store, cfg, x, is_training=True):
    """BigGAN.py:591-715, canonical flags.  Returns {"real": [B,1], "cls": [B,n_labels]}."""
    _, opt = make_opts(cfg, is_training, store)
    outputs = {}
    with store.variable_scope("discriminator"):
        block_info = d_block_info(cfg.img_size)
        ch = round_up(int(cfg.ch * 2.0 ** 0), 8)                                # :605, :588-589
        b_i = 0
        ch_mul = 1
        for block_count in block_info["counts"]:
            scope = "resblock_down_" + str(ch_mul)                              # :625
            for sb_i in range(block_count):
                if block_count > 1:
                    scope = scope + "_" + str(sb_i)
                x = R.resblock_down(store, x, channels=ch, use_bias=False, opt=opt, scope=scope)  # :633
            b_i += 1
            if b_i == block_info["sa_index"]:
                x = _sa2(store, cfg, x, ch, opt)                                # :636-637
            ch = round_up(int(cfg.ch * 2.0 ** b_i), 8)                          # :663
            ch_mul = ch_mul * 2
        ch = round_up(int(cfg.ch * 2.0 ** (b_i - 1)), 8)                        # :666
        x = R.resblock(store, x, channels=ch, use_bias=False, opt=opt, scope="resblock")   # :668
        x = opt["act"](x)                                                       # :669
        features = R.global_sum_pooling(x)                                      # :671
        x = R.fully_connected(store, features, units=1, opt=opt, scope="D_logit")          # :681-684
        outputs["real"] = x
        if cfg.acgan:
            cls_opt = copy.copy(opt)
            cls_opt["conv"] = copy.copy(opt["conv"])
            cls_opt["sn"] = False
            cls_opt["conv"]["sn"] = False                                       # :687, :1482-1487
            outputs["cls"] = R.fully_connected(store, features, units=cfg.n_labels, opt=cls_opt,
                                               scope="DC_logit")                # :700
    return outputs


# --------------------------------------------------------------------------------------------
# One training iteration: D update then G update (BigGAN.py:1061-1084)
# --------------------------------------------------------------------------------------------

class AdamState:
    """tf.train.AdamOptimizer (eps 1e-8 outside the bias correction) + optional
    MovingAverageOptimizer shadows (BigGAN.py:923-926)."""

    def __init__(self, lr, beta1, beta2, ema_decay=None):
        self.lr, self.b1, self.b2, self.ema_decay = lr, beta1, beta2, ema_decay
        self.t = 0
        self.m, self.v, self.shadow = {}, {}, {}

    def apply(self, store, grads):
        self.t += 1
        lr_t = self.lr * np.sqrt(1 - self.b2 ** self.t) / (1 - self.b1 ** self.t)
        for name, g in grads.items():
            p = store.vars[name]
            if name not in self.m:
                self.m[name] = torch.zeros_like(p)
                self.v[name] = torch.zeros_like(p)
                if self.ema_decay is not None:
                    self.shadow[name] = p.detach().clone()
            self.m[name] = self.b1 * self.m[name] + (1 - self.b1) * g
            self.v[name] = self.b2 * self.v[name] + (1 - self.b2) * g * g
            new_p = p.detach() - lr_t * self.m[name] / (torch.sqrt(self.v[name]) + 1e-8)
            if self.ema_decay is not None:
                d = self.ema_decay
                self.shadow[name] = d * self.shadow[name] + (1 - d) * new_p
            store.vars[name] = new_p.clone()


def _detach_all(store):
    for k in list(store.vars.keys()):
        store.vars[k] = store.vars[k].detach().clone()


def d_update(store, cfg, real, labels, z, cls_z, rnd_real=None, rnd_fake=None, opt_state=None,
             apply=True):
    """One d_train_op run (BigGAN.py:1063): returns (d_loss, grads dict)."""
    store.begin_phase()
    _detach_all(store)
    dt = store.dtype
    real = torch.as_tensor(real, dtype=dt)
    aug = R.DiffAugment(real, policy=cfg.da_policy, rnd=rnd_real)               # :806
    d_real = discriminator(store, cfg, aug)                                     # :807
    fake = generator(store, cfg, torch.as_tensor(z, dtype=dt), torch.as_tensor(cls_z, dtype=dt))  # :883
    d_fake = discriminator(store, cfg, R.DiffAugment(fake, policy=cfg.da_policy, rnd=rnd_fake))   # :857
    d_cls = 0
    if cfg.acgan:
        d_cls = cfg.d_cls_loss_weight * R.cls_loss_logistic(torch.as_tensor(labels, dtype=dt), d_real["cls"])  # :853
    d_loss = R.discriminator_loss(cfg.gan_type, real=d_real["real"], fake=d_fake["real"],
                                  flood_level=cfg.d_flood) + d_cls              # :879
    d_vars = store.trainables("discriminator")                                  # :916
    names = list(d_vars.keys())
    grads = torch.autograd.grad(d_loss, [d_vars[n] for n in names], allow_unused=True)
    gd = {n: (g if g is not None else torch.zeros_like(d_vars[n])) for n, g in zip(names, grads)}
    if apply and opt_state is not None:
        opt_state.apply(store, gd)
    return float(d_loss.detach()), gd, {"fake": fake.detach(), "real_logits": d_real["real"].detach(),
                                         "fake_logits": d_fake["real"].detach()}


def g_update(store, cfg, z, cls_z, rnd_fake=None, opt_state=None, apply=True):
    """One g_train_op run (BigGAN.py:1082): hinge G loss needs only D(G(z))."""
    store.begin_phase()
    _detach_all(store)
    dt = store.dtype
    cls_t = torch.as_tensor(cls_z, dtype=dt)
    fake = generator(store, cfg, torch.as_tensor(z, dtype=dt), cls_t)
    d_fake = discriminator(store, cfg, R.DiffAugment(fake, policy=cfg.da_policy, rnd=rnd_fake))
    g_cls = 0
    if cfg.acgan:
        g_cls = cfg.g_cls_loss_weight * R.cls_loss_logistic(cls_t, d_fake["cls"])   # :894
    g_loss = R.generator_loss(cfg.gan_type, fake=d_fake["real"], real=None,
                              flood_level=cfg.g_flood) + g_cls                   # :896
    if cfg.g_regularization != "none":
        for rl in store.regularization_losses():                                # :897-898 tf.add_n([g_loss] + reg losses)
            g_loss = g_loss + rl
    g_vars = store.trainables("generator")                                      # :917
    names = list(g_vars.keys())
    grads = torch.autograd.grad(g_loss, [g_vars[n] for n in names], allow_unused=True)
    gd = {n: (g if g is not None else torch.zeros_like(g_vars[n])) for n, g in zip(names, grads)}
    if apply and opt_state is not None:
        opt_state.apply(store, gd)
    return float(g_loss.detach()), gd, {"fake": fake.detach(), "fake_logits": d_fake["real"].detach()}


def synthetic_batch(cfg, B=None, seed=0):
    """SURVEY 8(d) synthetic inputs: real U(-1,1) seed 0, labels seed 1, z trunc-normal seeds 2/3."""
    B = B or cfg.batch_size
    S = cfg.img_size
    real = np.random.Generator(np.random.PCG64(seed + 0)).uniform(-1, 1, (B, S, S, cfg.c_dim)).astype(np.float32)
    lab = np.random.Generator(np.random.PCG64(seed + 1)).integers(0, max(cfg.n_labels, 1), B)
    labels = np.zeros((B, max(cfg.n_labels, 1)), np.float32)
    labels[np.arange(B), lab] = 1.0

    def tn(s):
        r = np.random.Generator(np.random.PCG64(s))
        x = r.standard_normal((B, 1, 1, cfg.z_dim))
        bad = np.abs(x) > 2
        while bad.any():
            x[bad] = r.standard_normal(int(bad.sum()))
            bad = np.abs(x) > 2
        return x.astype(np.float32)

    return {"real": real, "labels": labels, "z_d": tn(seed + 2), "z_g": tn(seed + 3)}
