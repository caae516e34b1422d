"""Generator / discriminator wiring of the reference for the canonical flag set:
BigGAN.generator (BigGAN.py:246-582) and BigGAN.discriminator (BigGAN.py:591-715).  This is plumbing
(z split, block loop, attention placement); all math lives in ops.py -> libbiggan_b200.so.
"""
import copy
from dataclasses import dataclass

import torch

from . import _lib
from .ops import (bn, cast, compute_dtype, rgb_pad, constant_initializer, conv, fully_connected, get_variable,
                  global_sum_pooling, lrelu, precompute_cbn_projections, prelu, relu, reshape, resblock, resblock_down,
                  resblock_up_condition, round_up, self_attention_2, variable_scope)
from .utils import orthogonal_regularizer, orthogonal_regularizer_fc
from .tensor import F32, Tensor


@dataclass
class Config:
    img_size: int = 128
    ch: int = 96
    batch_size: int = 256          # GLOBAL batch (split across data-parallel ranks)
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
    sa_gamma_init: float = 0.0     # ops.py:486 initialises to 0; parity tests use 0.1
    # SURVEY 8(f) "next" rows: the fork's CLI-default extras and layer variants (main.py:49-67)
    activation: str = "relu"                    # relu | prelu | lrelu            (BigGAN.py:71-84)
    upsampling_method: str = "deconv4"          # ops.py:200-218
    downsampling_method: str = "strided_conv3"  # ops.py:269-291
    g_conv: str = "deconv3"                     # ops.py:220-230
    bn_in_d: bool = False
    g_regularization: str = "none"              # none | ortho | ortho_cosine    (BigGAN.py:257-271)
    g_regularization_factor: float = 1e-4
    virtual_batches: int = 1                    # gradient accumulation, utils.py:242-267

    @property
    def depth(self):
        return self.img_size.bit_length() - 2                                   # BigGAN.py:19

    @property
    def acgan(self):
        return self.n_labels > 0


CONFIGS = {   # BASELINE.json configs 0-4
    0: dict(img_size=32, ch=32, batch_size=64, n_labels=10),
    1: dict(img_size=128, ch=96, batch_size=256, n_labels=1000),
    2: dict(img_size=128, ch=96, batch_size=128, n_labels=1000, da_policy="color,translation,cutout"),
    3: dict(img_size=256, ch=96, batch_size=256, n_labels=1000),
    4: dict(img_size=512, ch=64, batch_size=64, n_labels=1000),
}

_G_BLOCKS = {      # BigGAN.py:290-293 (+ 32x32 extrapolation)
    32: {"counts": [1, 1, 1], "sa_index": 2},
    64: {"counts": [1, 1, 1, 1], "sa_index": 3},
    128: {"counts": [1, 1, 1, 1, 1], "sa_index": 4},
    256: {"counts": [1, 2, 1, 1, 1], "sa_index": 3},
    512: {"counts": [1, 2, 1, 1, 2], "sa_index": 3},
}
_D_BLOCKS = {      # BigGAN.py:607-610
    32: {"counts": [1, 1, 1], "sa_index": 1},
    64: {"counts": [1, 1, 1, 1], "sa_index": 1},
    128: {"counts": [1, 1, 1, 1, 1], "sa_index": 1},
    256: {"counts": [1, 1, 1, 2, 1], "sa_index": 2},
    512: {"counts": [1, 2, 1, 1, 2], "sa_index": 2},
}


def _lrelu_02(x):
    return lrelu(x, 0.2)


def activation_fn(cfg):
    """BigGAN.py:71-84."""
    if cfg.activation == 'relu':
        return relu
    if cfg.activation == 'prelu':
        return prelu
    if cfg.activation == 'lrelu':
        return _lrelu_02
    raise ValueError("Unknown activation function: " + str(cfg.activation))


def _opts(cfg, is_training):
    bn_options = {"type": "batch_norm", "momentum": cfg.bn_momentum}            # GANBase.py:39-41
    conv_options = {"padding_type": "reflect", "sn": True, "regularizer": None}  # GANBase.py:49-51
    act = activation_fn(cfg)
    g_opt = {"sn": True, "is_training": is_training, "upsampling_method": cfg.upsampling_method, "g_conv": cfg.g_conv,
             "act": act, "self_attention_bias": True, "bn": copy.deepcopy(bn_options),
             "conv": copy.deepcopy(conv_options), "fc_regularizer": None}        # BigGAN.py:248-274
    if is_training and cfg.g_regularization != 'none':                          # BigGAN.py:257-271
        if cfg.g_regularization not in ('ortho', 'ortho_cosine'):
            raise ValueError("Unknown regularization method: " + str(cfg.g_regularization))
        g_opt["conv"]["regularizer"] = orthogonal_regularizer(cfg.g_regularization_factor, type=cfg.g_regularization)
        g_opt["fc_regularizer"] = orthogonal_regularizer_fc(cfg.g_regularization_factor, type=cfg.g_regularization)
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


def _attention(cfg, x, ch, opt):
    with variable_scope("self_attention"):
        get_variable("gamma", [1], initializer=constant_initializer(cfg.sa_gamma_init))
    return self_attention_2(x, channels=ch, opt=opt, scope='self_attention')


def generator(cfg, z, cls_z, is_training=True):
    """BigGAN.py:246-582.  z: torch fp32 [B,1,1,z_dim] buffer; cls_z: torch fp32 [B,n_labels]."""
    opt, _ = _opts(cfg, is_training)
    if cfg.img_size not in _G_BLOCKS:
        raise ValueError("Invalid image size specified: " + str(cfg.img_size))
    with variable_scope("generator"):
        split_sizes = z_split_sizes(cfg)
        block_info = _G_BLOCKS[cfg.img_size]
        B = z.shape[0]
        z2 = z.reshape(B, cfg.z_dim)
        parts = list(torch.split(z2, split_sizes, dim=-1))                      # :335 (plumbing)
        if cfg.acgan:
            parts = [torch.cat([p, cls_z.reshape(B, cfg.n_labels)], dim=-1) for p in parts]   # :358-365
        z_split = [Tensor(p.contiguous()) for p in parts]
        n_blocks = len(block_info["counts"])
        ch_mul = 2 ** (n_blocks - 1)
        ch = round_up(int(cfg.ch * 2.0 ** (n_blocks - 1)), 8)                   # :428
        f_width = round_up((split_sizes[0] + cfg.n_labels) * 1.85, 8)           # :433
        if opt["act"] is relu:                                                  # :434 (no scope, backward compatibility)
            x = fully_connected(z_split[0], units=f_width, scope='dense1', opt=opt)  # :436
            x = relu(x)
            x = fully_connected(x, units=4 * 4 * ch, scope='dense2', opt=opt)   # :438
        else:
            with variable_scope('first'):                                       # :440-443
                x = fully_connected(z_split[0], units=f_width, scope='dense1', opt=opt)
                x = opt["act"](x)
                x = fully_connected(x, units=4 * 4 * ch, scope='dense2', opt=opt)
        x = reshape(x, [-1, 4, 4, ch])                                          # :446
        x = cast(x, compute_dtype())
        # every conditional-BN gamma / beta projection of the pass in one grouped launch (same names, same math; ops.py)
        hoist, zi, cin, ch_h, mul_h = [], 1, ch, ch, ch_mul
        for bi_h, block_count in enumerate(block_info["counts"]):
            scope = 'resblock_up_' + str(mul_h)
            for sb_i in range(block_count):
                if block_count > 1:
                    scope = scope + '_' + str(sb_i)
                hoist += [(scope + '/res1/batch_norm', z_split[zi], cin), (scope + '/res2/batch_norm', z_split[zi], ch_h)]
                zi, cin = zi + 1, ch_h
            ch_h = round_up(int(cfg.ch * 2.0 ** (n_blocks - bi_h - 2)), 8)
            mul_h //= 2
        precompute_cbn_projections(hoist, opt)
        zi = 1
        b_i = 0
        for block_count in block_info["counts"]:
            scope = 'resblock_up_' + str(ch_mul)
            for sb_i in range(block_count):
                layer_z = z_split[zi]
                zi += 1
                if block_count > 1:
                    scope = scope + '_' + str(sb_i)
                x = resblock_up_condition(x, layer_z, channels=ch, use_bias=False, opt=opt, scope=scope)   # :479
            b_i += 1
            if b_i == block_info["sa_index"]:
                x = _attention(cfg, x, ch, opt)                                 # :482-483
            ch = round_up(int(cfg.ch * 2.0 ** (n_blocks - b_i - 1)), 8)
            ch_mul = ch_mul // 2
        if opt["act"] is relu:
            x = bn(x, opt=opt, _act=_lib.ACT_RELU)                              # :491-492 (BN + act fused)
        else:
            x = opt["act"](bn(x, opt=opt))
        # bf16 tensor-core mode: the image is produced with zero-padded channels (3 -> 16) so that G_logit and
        # D's first convs are tcgen05 GEMMs; `ops.slice_channels(x, 3)` gives the plain RGB tensor
        x = conv(x, channels=cfg.c_dim, kernel=3, stride=1, pad=1, use_bias=False, opt=opt, scope='G_logit',
                 _act=_lib.ACT_TANH, _pad_cout_to=rgb_pad())                    # :570 + tanh :580
        return x


def discriminator(cfg, x, is_training=True):
    """BigGAN.py:591-715.  Returns {"real": [B,1], "cls": [B,n_labels]} (fp32)."""
    _, opt = _opts(cfg, is_training)
    if cfg.img_size not in _D_BLOCKS:
        raise ValueError("Invalid image size specified: " + str(cfg.img_size))
    outputs = {}
    with variable_scope("discriminator"):
        block_info = _D_BLOCKS[cfg.img_size]
        ch = round_up(int(cfg.ch), 8)
        b_i = 0
        ch_mul = 1
        for block_count in block_info["counts"]:
            scope = 'resblock_down_' + str(ch_mul)
            for sb_i in range(block_count):
                if block_count > 1:
                    scope = scope + '_' + str(sb_i)
                x = resblock_down(x, channels=ch, use_bias=False, opt=opt, scope=scope)   # :633
            b_i += 1
            if b_i == block_info["sa_index"]:
                x = _attention(cfg, x, ch, opt)                                 # :636-637
            ch = round_up(int(cfg.ch * 2.0 ** b_i), 8)
            ch_mul = ch_mul * 2
        ch = round_up(int(cfg.ch * 2.0 ** (b_i - 1)), 8)                        # :666
        x = resblock(x, channels=ch, use_bias=False, opt=opt, scope='resblock')  # :668
        if opt["act"] is relu:
            features = global_sum_pooling(x, _relu_first=True)                  # :669-671
        else:
            features = global_sum_pooling(opt["act"](x))
        features = cast(features, F32)
        x = fully_connected(features, units=1, opt=opt, scope='D_logit')        # :681-684
        outputs["real"] = x
        if cfg.acgan:
            cls_opt = copy.copy(opt)
            cls_opt["conv"] = copy.copy(opt["conv"])
            cls_opt["sn"] = False
            cls_opt["conv"]["sn"] = False                                       # :687, :1482-1487
            outputs["cls"] = fully_connected(features, units=cfg.n_labels, opt=cls_opt, scope='DC_logit')  # :700
    return outputs
