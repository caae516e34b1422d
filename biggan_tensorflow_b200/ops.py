"""Operator surface of the reference's ops.py, re-implemented on libbiggan_b200.so.

Same function names, positional order and defaults as /root/reference/ops.py (cited per function),
operating on NHWC device tensors (`tensor.Tensor`).  Variables are created through the
`tf_shim.variable_scope/get_variable` shim with the reference's names.  There is no TF and no CPU
fallback: every computation is a call into the C ABI (include/biggan_b200.h).

Precision modes (north_star): `set_compute_dtype(F32)` = parity mode (fp32 SIMT kernels),
`set_compute_dtype(BF16)` = performance mode (bf16 activations, fp32 accumulate, tcgen05 kernels
where the shape is eligible).  Fully-connected layers always run in fp32 (they are tiny).
"""
import contextlib
import ctypes
import os

import torch

from . import _lib
from ._lib import ConvDesc, call, query
from .tensor import (BF16, F32, Tensor, accumulate_grad, bg_dtype, empty, grad_target, record,
                     recording, stream, workspace)
from .tf_shim import (constant_initializer, current_scope_name, get_store, get_variable,
                      random_normal_initializer, truncated_normal_initializer, variable_scope)

##################################################################################
# Initialization (ops.py:13-14)
##################################################################################

weight_init = truncated_normal_initializer(mean=0.0, stddev=0.02)
gan_dtype = torch.float32

_state = {"compute_dtype": F32, "force_simt": 0}


def set_compute_dtype(dt):
    _state["compute_dtype"] = dt


def compute_dtype():
    return _state["compute_dtype"]


def set_force_simt(flag):
    _state["force_simt"] = 1 if flag else 0


class _Dist:
    """Data-parallel hooks: cross-replica BN statistics + global-batch loss means."""
    world_size = 1

    @staticmethod
    def all_reduce_sum(buf):
        return buf


dist = _Dist()


def round_up(val, multiple):
    return (int(val) + multiple - 1) // multiple * multiple


##################################################################################
# plumbing ops (views / casts / adds)
##################################################################################

def reshape(x, shape):
    shape = [int(s) for s in shape]
    y = Tensor(x.data.reshape(shape), requires_grad=x.requires_grad)
    if recording(x):
        def bw():
            if y.grad is not None:
                accumulate_grad(x, y.grad.reshape(x.data.shape), own=False)
        record(bw)
    return y


def cast(x, dtype):
    if x.dtype == dtype:
        return x
    y = Tensor(empty(x.shape, dtype), requires_grad=x.requires_grad)
    call("bg_cast", x.data, bg_dtype(x), y.data, bg_dtype(y), x.numel(), stream())
    if recording(x):
        def bw():
            if y.grad is None:
                return
            g = empty(x.shape, x.dtype)
            call("bg_cast", y.grad, bg_dtype(y.grad), g, bg_dtype(g), g.numel(), stream())
            accumulate_grad(x, g, own=True)
        record(bw)
    return y


def add(a, b):
    """x + x_init of ops.py:198/266/313 (and scalar loss sums)."""
    y = Tensor(empty(a.shape, a.dtype), requires_grad=a.requires_grad or b.requires_grad)
    call("bg_add", a.data, b.data, y.data, y.numel(), bg_dtype(y), stream())
    if recording(a, b):
        def bw():
            if y.grad is None:
                return
            if a.requires_grad:
                accumulate_grad(a, y.grad, own=False)
            if b.requires_grad:
                accumulate_grad(b, y.grad, own=False)
        record(bw)
    return y


def scale(x, alpha):
    """python-float * tensor (e.g. d_cls_loss_weight * cls_loss, BigGAN.py:853)."""
    y = Tensor(empty(x.shape, x.dtype), requires_grad=x.requires_grad)
    call("bg_axpby", x.data, None, None, float(alpha), 0.0, y.data, y.numel(), bg_dtype(y), stream())
    if recording(x):
        def bw():
            if y.grad is None:
                return
            g = torch.empty_like(y.grad)
            call("bg_axpby", y.grad, None, None, float(alpha), 0.0, g, g.numel(), bg_dtype(g), stream())
            accumulate_grad(x, g, own=True)
        record(bw)
    return y


Tensor.__add__ = lambda a, b: add(a, b) if isinstance(b, Tensor) else (a if b == 0 else NotImplemented)
Tensor.__radd__ = lambda a, b: a if b == 0 else NotImplemented
Tensor.__mul__ = lambda a, b: scale(a, b)
Tensor.__rmul__ = lambda a, b: scale(a, b)


##################################################################################
# Spectral norm (ops.py:718-747)
##################################################################################

def spectral_norm(w, iteration=1, _out_dtype=None, _taps=1, _want_t=False, _pad_c=0, _pad_r=0):
    """w: Variable [..., C].  Returns W/sigma (same shape) after one power-iteration step; the
    persistent `u` (created in the current scope, ops.py:722) is advanced in place (ops.py:743).
    Evaluated once per update phase and variable (SURVEY 8a row 3 state semantics)."""
    if iteration != 1:
        raise ValueError("only iteration=1 is supported")
    store = get_store()
    w_shape = list(w.shape)
    C = w_shape[-1]
    K = w.numel() // C
    u = get_variable("u", [1, C], initializer=random_normal_initializer(), trainable=False)
    out_dtype = _out_dtype or F32
    c_pad = max(int(_pad_c), C)      # zero-padded output channels (tensor-core eligibility of narrow convs)
    rows = K // _taps
    r_pad = max(int(_pad_r), rows)   # zero-padded input channels per tap (RGB layers)
    key = (u.name, out_dtype, bool(_want_t), c_pad, r_pad)
    hit = store.sn_cache.get(key)
    if hit is not None:
        return hit
    grouped = getattr(store, "sn_grouped", {}).get(u.name)
    if grouped is not None:
        # the grouped pass (sn_group.SnGroup.forward) has advanced this u for the phase; evaluating the weight again in
        # another layout would advance it a second time (the reference updates u once per run, ops.py:743)
        raise RuntimeError("spectral_norm(%s): layout %r was not part of the grouped pass built for this network (it holds %r); "
                           "rebuild the Trainer with the options in use" % (u.name, key[1:], grouped[1:]))
    reg = getattr(store, "sn_registry", None)
    if reg is not None and key not in reg:       # the Trainer is enumerating the network's spectral norms
        reg[key] = {"w": w, "u": u, "taps": _taps, "rows": rows, "r_pad": r_pad, "c_pad": c_pad, "out_dtype": out_dtype,
                    "want_t": bool(_want_t), "key": key}
    base = store.sn_cache.get(("pi", u.name))
    if base is None:
        v_hat = empty([K], F32)
        sigma = empty([1], F32)
        ws = workspace(query("bg_sn_workspace_bytes", K, C))
        call("bg_sn_power_iter", w.data, u.data, u.data, v_hat, sigma, K, C, ws, ws.numel(), stream())
        base = (v_hat, sigma)
        store.sn_cache[("pi", u.name)] = base
    v_hat, sigma = base
    p_shape = w_shape[:-2] + [r_pad, c_pad]
    alloc = empty if (c_pad == C and r_pad == rows) else (lambda sh, dt: torch.zeros(tuple(sh), dtype=dt, device=w.data.device))
    wn = Tensor(alloc(p_shape, out_dtype), requires_grad=w.requires_grad)
    w_t = alloc(p_shape, out_dtype) if _want_t else None
    call("bg_sn_scale", w.data, sigma, wn.data, w_t, bg_dtype(out_dtype), _taps, rows, C, r_pad, c_pad, stream())
    wn.aux = {"w_t": w_t, "sigma": sigma, "v_hat": v_hat, "u_hat": u.data}
    store.sn_cache[key] = wn
    if recording(w):
        def bw():
            if wn.grad is None:
                return
            dst, fin = grad_target(w)
            ws2 = workspace(query("bg_sn_workspace_bytes", K, C))
            call("bg_sn_bwd", w.data, wn.grad, rows, r_pad, c_pad, u.data, v_hat, sigma, dst, K, C, ws2, ws2.numel(), stream())
            if fin:
                fin()
        record(bw)
    return wn


##################################################################################
# Convolution family (ops.py:49-139, :148-175)
##################################################################################

def _conv_flops(d):
    """Algorithmic FLOPs: conv 2 k^2 Cin Cout Hout Wout, deconv 2 k^2 Cin Cout Hin Win."""
    pix = d.B * (d.Hi * d.Wi if d.transposed else d.Ho * d.Wo)
    return 2.0 * d.kh * d.kw * d.Ci * d.Co * pix


def _conv_tag(d):
    if d.Hi == 1 and d.Wi == 1 and d.kh == 1:
        return "[fc]"
    return "[%s k%d s%d %dx%d %d>%d b%d]" % ("deconv" if d.transposed else "conv", d.kh, d.stride, d.Hi, d.Wi, d.Ci, d.Co, d.B)


def _conv_op(x, wn, bias, desc, act=_lib.ACT_NONE, residual=None, pre_relu=False):
    """y = act(conv(x; wn) + bias + residual) and its tape entry (dgrad / wgrad / bias grad).
    pre_relu (bf16 reflect-padded convs only): y = conv(relu(x)); the ReLU rides in the padding copy."""
    dt = bg_dtype(x)
    fs = _state["force_simt"]
    y = Tensor(empty([desc.B, desc.Ho, desc.Wo, desc.Co], x.dtype),
               requires_grad=x.requires_grad or wn.requires_grad or (bias is not None and bias.requires_grad)
               or (residual is not None and residual.requires_grad))
    w_t = wn.aux.get("w_t") if wn.aux else None
    # bf16 mode, reflect padding: pad ONCE into a tensor that lives until the backward pass; forward and weight gradient
    # then run as plain (pad-free) convs on it instead of each re-padding x into transient workspace
    xin, desc_f = x.data, desc
    if (x.dtype == BF16 and not desc.transposed and desc.pad_mode == _lib.PAD_REFLECT
            and (desc.pad_t | desc.pad_l | desc.pad_b | desc.pad_r)):
        Hp, Wp = desc.Hi + desc.pad_t + desc.pad_b, desc.Wi + desc.pad_l + desc.pad_r
        xin = empty([desc.B, Hp, Wp, desc.Ci], x.dtype)
        call("bg_reflect_pad_relu" if pre_relu else "bg_reflect_pad", x.data, xin, desc.B, desc.Hi, desc.Wi, desc.Ci, desc.pad_t,
             desc.pad_l, desc.pad_b, desc.pad_r, dt, stream(), tag="reflect_pad")
        desc_f = ConvDesc(desc.B, Hp, Wp, desc.Ci, desc.Ho, desc.Wo, desc.Co, desc.kh, desc.kw, desc.stride, 0, 0, 0, 0,
                          _lib.PAD_ZERO, 0)
    else:
        assert not pre_relu, "pre_relu needs the materialised reflect padding (bf16 mode)"
    nbf = query("bg_conv2d_fwd_workspace_bytes", desc_f, dt)
    wsf = workspace(nbf) if nbf else None
    call("bg_conv2d_fwd", desc_f, xin, wn.data, w_t, bias.data if bias is not None else None,
         residual.data if residual is not None else None, y.data, act, dt, wsf, wsf.numel() if nbf else 0, fs, stream(),
         flops=_conv_flops(desc), tag="conv_fwd" + _conv_tag(desc))
    y.relu_out = act == _lib.ACT_RELU
    # the gradient w.r.t. x passes a ReLU whose mask is (x.data > 0) -- x is a fused-ReLU output, or this conv applies the ReLU in
    # its padding copy: the reflect-padded dgrad applies that mask where it produces the gradient (bg_conv2d_dgrad relu_mask)
    reflect = (not desc.transposed and desc.pad_mode == _lib.PAD_REFLECT and (desc.pad_t | desc.pad_l | desc.pad_b | desc.pad_r) != 0)
    mask_x = reflect and (pre_relu or x.relu_out) and os.environ.get("BG_NO_RELU_MASK_FUSION") != "1"
    if recording(x, wn, bias, residual):
        def bw():
            dy = y.grad
            if dy is None:
                return
            if act == _lib.ACT_TANH:
                g = torch.empty_like(dy)
                call("bg_tanh_bwd", y.data, dy, g, g.numel(), bg_dtype(g), stream())
                dy = g
            elif act == _lib.ACT_RELU and not (y.n_contribs > 0 and y.n_masked == y.n_contribs):
                g = torch.empty_like(dy)      # (skipped when every consumer's dgrad already applied this mask)
                call("bg_relu_bwd", y.data, dy, g, g.numel(), bg_dtype(g), stream())
                dy = g
            if residual is not None and residual.requires_grad:
                accumulate_grad(residual, dy, own=False)
            if wn.requires_grad:
                dst, fin = grad_target(wn, dtype=F32)          # weight gradients are fp32 in every mode
                nb = query("bg_conv2d_wgrad_workspace_bytes", desc_f, dt)
                ws = workspace(nb)
                call("bg_conv2d_wgrad", desc_f, xin, dy, dst, dt, ws, ws.numel(), fs, stream(),
                     flops=_conv_flops(desc), tag="conv_wgrad" + _conv_tag(desc))
                if fin:
                    fin()
            if bias is not None and bias.requires_grad:
                rows = desc.B * desc.Ho * desc.Wo
                dst, fin = grad_target(bias)
                ws = workspace(query("bg_colsum_workspace_bytes", rows, desc.Co))
                call("bg_colsum", dy, dst, rows, desc.Co, dt, ws, ws.numel(), stream())
                if fin:
                    fin()
            if x.requires_grad:
                nb = query("bg_conv2d_dgrad_workspace_bytes", desc, dt)
                ws = workspace(nb) if nb else None
                mask = x.data if mask_x else None
                if pre_relu and not mask_x:              # (fusion switched off) d relu(x) first, then the mask in its own pass
                    dxr = torch.empty_like(x.data)
                    call("bg_conv2d_dgrad", desc, dy, wn.data, w_t, None, None, dxr, dt, ws, ws.numel() if nb else 0, fs, stream(),
                         flops=_conv_flops(desc), tag="conv_dgrad" + _conv_tag(desc))
                    dst, fin = grad_target(x)
                    call("bg_relu_bwd", x.data, dxr, dst, x.numel(), dt, stream())
                    if fin:
                        fin()
                    return
                res = x.grad
                if res is None:
                    dx = torch.empty_like(x.data)
                elif x.grad_owned:
                    dx = res                      # fused accumulate in the dgrad epilogue
                else:
                    dx = torch.empty_like(x.data)
                call("bg_conv2d_dgrad", desc, dy, wn.data, w_t, res, mask, dx, dt, ws, ws.numel() if nb else 0, fs, stream(),
                     flops=_conv_flops(desc), tag="conv_dgrad" + _conv_tag(desc))
                x.n_contribs += 1
                if mask_x and x.relu_out:
                    x.n_masked += 1
                x.grad, x.grad_owned, x.grad_set = dx, True, True
        record(bw)
    return y


def _sn_weight(w, taps, conv_like, x_dtype, pad_c=0, pad_r=0):
    """Normalised weights in the layout(s) the conv kernels want."""
    if x_dtype == BF16:
        return spectral_norm(w, _out_dtype=BF16, _taps=taps, _want_t=True, _pad_c=pad_c, _pad_r=pad_r)
    return spectral_norm(w, _out_dtype=F32, _taps=taps, _pad_c=pad_c, _pad_r=pad_r)


def rgb_pad():
    """Channel count the 3-channel image tensors are zero-padded to in bf16 tensor-core mode (0 = no padding):
    with 16 channels the RGB convs (3 -> ch in D, ch -> 3 `G_logit`) run on the tcgen05 kernels."""
    if _state["compute_dtype"] == BF16 and not _state["force_simt"] and _lib.lib.bg_has_tensor_core_path() == 1:
        return 16
    return 0


def c_real(x):
    """Logical channel count of a (possibly zero-padded) NHWC tensor."""
    a = x.aux
    return a["c_real"] if isinstance(a, dict) and "c_real" in a else x.shape[-1]


def pad_channels(x, c_pad, dtype=None):
    """[.., C] -> [.., c_pad] with zero fill (and optional cast); the result remembers C as its logical width."""
    C = x.shape[-1]
    dtype = dtype or x.dtype
    y = Tensor(empty(list(x.shape[:-1]) + [c_pad], dtype), requires_grad=x.requires_grad)
    y.aux = {"c_real": C}
    npix = x.numel() // C
    call("bg_pad_channels", x.data, bg_dtype(x), y.data, bg_dtype(y), npix, C, c_pad, stream())
    if recording(x):
        def bw():
            if y.grad is None:
                return
            g = torch.empty_like(x.data)
            call("bg_pad_channels", y.grad, bg_dtype(y.grad), g, bg_dtype(g), npix, c_pad, C, stream())
            accumulate_grad(x, g, own=True)
        record(bw)
    return y


def slice_channels(x, C):
    """[.., Cp] -> [.., C] (first C channels)."""
    Cp = x.shape[-1]
    if Cp == C:
        return x
    y = Tensor(empty(list(x.shape[:-1]) + [C], x.dtype), requires_grad=x.requires_grad)
    npix = x.numel() // Cp
    call("bg_pad_channels", x.data, bg_dtype(x), y.data, bg_dtype(y), npix, Cp, C, stream())
    if recording(x):
        def bw():
            if y.grad is None:
                return
            g = torch.empty_like(x.data)
            call("bg_pad_channels", y.grad, bg_dtype(y.grad), g, bg_dtype(g), npix, C, Cp, stream())
            accumulate_grad(x, g, own=True)
        record(bw)
    return y


def _padded_bias(bias, c_pad):
    """Zero-padded copy of a bias vector (plumbing) whose gradient flows back to the first C entries."""
    C = bias.shape[0]
    pb = Tensor(torch.zeros(c_pad, dtype=F32, device=bias.data.device), requires_grad=bias.requires_grad)
    pb.data[:C].copy_(bias.data)
    if recording(bias):
        def bw():
            if pb.grad is not None:
                accumulate_grad(bias, pb.grad[:C].contiguous(), own=True)
        record(bw)
    return pb


def _padded_param(p, c_pad, fill):
    """Copy of a per-channel parameter [C] extended to c_pad entries with `fill` (plumbing); the gradient of the first C
    entries flows back."""
    C = p.shape[0]
    pp = Tensor(torch.full((c_pad,), float(fill), dtype=F32, device=p.data.device), requires_grad=p.requires_grad)
    pp.data[:C].copy_(p.data)
    if recording(p):
        def bw():
            if pp.grad is not None:
                accumulate_grad(p, pp.grad[:C].contiguous(), own=True)
        record(bw)
    return pp


def _conv_geometry(x, channels, kernel, stride, pad, opt):
    """pad logic of ops.py:65-84."""
    B, H, W, Cin = x.shape
    pt = pb = pl = pr = 0
    mode = _lib.PAD_ZERO
    if pad > 0:
        pad_type = opt.get("conv", {}).get("padding_type", "reflect")
        if H % stride == 0:
            pad = pad * 2
        else:
            pad = max(kernel - (H % stride), 0)
        pt = int(pad // 2)
        pb = int(pad - pt)
        pl = int(pad // 2)
        pr = int(pad - pl)
        if pad_type == "zero":
            # TF 'SAME' on the unpadded input (ops.py:80); pads follow the SAME rule
            def same(n):
                out = -(-n // stride)
                tot = max((out - 1) * stride + kernel - n, 0)
                return tot // 2, tot - tot // 2
            pt, pb = same(H)
            pl, pr = same(W)
            mode = _lib.PAD_ZERO
        elif pad_type == "reflect":
            mode = _lib.PAD_REFLECT
        else:
            raise ValueError("Unsupported padding type: " + str(pad_type))
    Ho = (H + pt + pb - kernel) // stride + 1
    Wo = (W + pl + pr - kernel) // stride + 1
    return ConvDesc(B, H, W, Cin, Ho, Wo, channels, kernel, kernel, stride, pt, pl, pb, pr, mode, 0)


def _im2col_rgb(x, desc, cz):
    """[B,H,W,Cx] (3 real channels) -> [B,Ho,Wo,cz] patches (bg_im2col_rgb); backward = its adjoint (bg_col2im_rgb)."""
    z = Tensor(empty([desc.B, desc.Ho, desc.Wo, cz], x.dtype), requires_grad=x.requires_grad)
    args = (desc.B, desc.Hi, desc.Wi, x.shape[-1], desc.Ho, desc.Wo, desc.kh, desc.kw, desc.stride)
    call("bg_im2col_rgb", x.data, z.data, *args, desc.pad_t, desc.pad_l, desc.pad_mode, cz, stream(), tag="im2col_rgb")
    if recording(x):
        def bw():
            if z.grad is None:
                return
            g = torch.empty_like(x.data)
            call("bg_col2im_rgb", z.grad, g, *args, desc.pad_t, desc.pad_l, desc.pad_b, desc.pad_r, desc.pad_mode, cz, stream(),
                 tag="col2im_rgb")
            accumulate_grad(x, g, own=True)
        record(bw)
    return z


def _conv_rgb_out(x, wn, desc, c_pad, act):
    """3x3 / stride 1 / one-pixel-padded conv to 3 (zero-padded to c_pad) channels in tensor-core mode -- `G_logit`, BigGAN.py:570.
    The nine taps are stacked along N: ONE 1x1 GEMM Ci -> 27 (32) on the unpadded input, then the shifted tap sum + activation
    (bg_rgb_tapsum_fwd).  Backward: the tap sum's adjoint, then a 1x1 dgrad (with the ReLU mask of x) and a 1x1 wgrad.
    A 3x3 conv with N = 16 costs 54 tensor-core instructions per 128-pixel tile at the operand-read floor of the tensor pipe
    (7x the layer's HBM time) plus a reflect-padded copy of x, a padded-domain gradient and a fold pass."""
    dt = _lib.BG_BF16
    fs = _state["force_simt"]
    B, H, W, Ci = desc.B, desc.Hi, desc.Wi, desc.Ci
    wall, wall_t = empty([Ci, 32], BF16), empty([32, Ci], BF16)
    call("bg_rgb_stack_weights", wn.data, wall, wall_t, Ci, c_pad, stream())
    d1 = ConvDesc(B, H, W, Ci, H, W, 32, 1, 1, 1, 0, 0, 0, 0, _lib.PAD_ZERO, 0)
    ycol = empty([B, H, W, 32], BF16)
    nbf = query("bg_conv2d_fwd_workspace_bytes", d1, dt)
    wsf = workspace(nbf) if nbf else None
    call("bg_conv2d_fwd", d1, x.data, wall, wall_t, None, None, ycol, _lib.ACT_NONE, dt, wsf, wsf.numel() if nbf else 0, fs, stream(),
         flops=_conv_flops(desc), tag="conv_fwd" + _conv_tag(desc))
    y = Tensor(empty([B, H, W, c_pad], BF16), requires_grad=x.requires_grad or wn.requires_grad)
    call("bg_rgb_tapsum_fwd", ycol, None, y.data, B, H, W, c_pad, desc.pad_mode, act, stream(), tag="rgb_tapsum")
    y.relu_out = act == _lib.ACT_RELU
    mask_x = bool(x.relu_out) and os.environ.get("BG_NO_RELU_MASK_FUSION") != "1"
    if recording(x, wn):
        def bw():
            if y.grad is None:
                return
            dcol = empty([B, H, W, 32], BF16)
            call("bg_rgb_tapsum_bwd", y.data, y.grad, dcol, B, H, W, c_pad, desc.pad_mode, act, stream(), tag="rgb_tapsum")
            if wn.requires_grad:
                dwall = empty([Ci, 32], F32)
                nb = query("bg_conv2d_wgrad_workspace_bytes", d1, dt)
                ws = workspace(nb)
                call("bg_conv2d_wgrad", d1, x.data, dcol, dwall, dt, ws, ws.numel(), fs, stream(),
                     flops=_conv_flops(desc), tag="conv_wgrad" + _conv_tag(desc))
                dst, fin = grad_target(wn, dtype=F32)
                call("bg_rgb_unstack_wgrad", dwall, dst, Ci, c_pad, stream())
                if fin:
                    fin()
            if x.requires_grad:
                nb = query("bg_conv2d_dgrad_workspace_bytes", d1, dt)
                ws = workspace(nb) if nb else None
                res = x.grad
                dx = res if (res is not None and x.grad_owned) else torch.empty_like(x.data)
                if mask_x:
                    call("bg_conv2d_dgrad", d1, dcol, wall, wall_t, res, x.data, dx, dt, ws, ws.numel() if nb else 0, fs, stream(),
                         flops=_conv_flops(desc), tag="conv_dgrad" + _conv_tag(desc))
                    x.n_masked += 1
                else:
                    call("bg_conv2d_dgrad", d1, dcol, wall, wall_t, res, None, dx, dt, ws, ws.numel() if nb else 0, fs, stream(),
                         flops=_conv_flops(desc), tag="conv_dgrad" + _conv_tag(desc))
                x.n_contribs += 1
                x.grad, x.grad_owned, x.grad_set = dx, True, True
        record(bw)
    return y


def conv(x, channels, opt, kernel=4, stride=2, pad=0, dilation=1, use_bias=True, scope='conv_0',
         _residual=None, _act=_lib.ACT_NONE, _pad_cout_to=0, _pre_relu=False):
    """ops.py:49-113.  _pre_relu: conv(relu(x)) (the activation resblock_down applies in front of its first conv)."""
    if _pre_relu and not (x.dtype == BF16 and pad > 0 and opt.get("conv", {}).get("padding_type", "reflect") == "reflect"
                          and x.shape[-1] % 8 == 0 and c_real(x) == x.shape[-1] and not isinstance(kernel, str)):
        x, _pre_relu = relu(x), False                # no padding copy to ride in: the plain activation
    with variable_scope(scope):
        if isinstance(kernel, str):
            raise ValueError("channel-sliced multi-kernel conv (ops.py:52-59) is out of scope")
        if dilation != 1:
            raise ValueError("dilation != 1 is not supported")
        desc = _conv_geometry(x, channels, kernel, stride, pad, opt)
        if not opt.get("conv", {}).get("sn", True):
            raise ValueError("conv with sn=False (tf.layers.conv2d path) is not supported")
        cin_x, cin = x.shape[-1], c_real(x)          # x may carry zero-padded channels (RGB tensors)
        reg = opt.get("conv", {}).get("regularizer", None) if 'generator' in current_scope_name() else None   # ops.py:87-92
        w = get_variable("kernel", shape=[kernel, kernel, cin, channels], initializer=weight_init, regularizer=reg)
        c_pad = round_up(channels, _pad_cout_to) if _pad_cout_to else channels
        if (cin == 3 and cin_x != cin and x.dtype == BF16 and c_pad % 16 == 0 and kernel * kernel * 3 <= 32
                and not _state["force_simt"] and _lib.lib.bg_has_tensor_core_path() == 1):
            # image input in tensor-core mode: im2col the 3 real channels (K = k*k*3 -> 32) and run ONE 1x1 GEMM instead of
            # k*k passes over 16 zero-padded channels; the HWIO kernel [k,k,3,Co] read as [27, Co] is already the GEMM operand
            z = _im2col_rgb(x, desc, 32)
            wn = _sn_weight(w, 1, True, x.dtype, pad_c=c_pad, pad_r=32)
            bias = None
            if use_bias:
                bias = get_variable("bias", [channels], initializer=constant_initializer(0.0))
                if c_pad != channels:
                    bias = _padded_bias(bias, c_pad)
            d1 = ConvDesc(desc.B, desc.Ho, desc.Wo, 32, desc.Ho, desc.Wo, c_pad, 1, 1, 1, 0, 0, 0, 0, _lib.PAD_ZERO, 0)
            y = _conv_op(z, wn, bias, d1, act=_act, residual=_residual)
            if c_pad != channels:
                y.aux = {"c_real": channels}
            return y
        wn = _sn_weight(w, kernel * kernel, True, x.dtype, pad_c=c_pad, pad_r=cin_x if cin_x != cin else 0)
        if (channels == 3 and c_pad % 8 == 0 and c_pad > 3 and kernel == 3 and stride == 1 and not use_bias and _residual is None
                and not _pre_relu and (desc.pad_t, desc.pad_l, desc.pad_b, desc.pad_r) == (1, 1, 1, 1) and x.dtype == BF16
                and cin_x == cin and cin % 16 == 0 and x.shape[1] >= 2 and x.shape[2] >= 2 and not _state["force_simt"]
                and _lib.lib.bg_has_tensor_core_path() == 1 and os.environ.get("BG_RGB_TAPSUM") != "0"):
            # image OUTPUT in tensor-core mode (G_logit): nine taps stacked along N, one 1x1 GEMM + shifted tap sum
            y = _conv_rgb_out(x, wn, desc, c_pad, _act)
            y.aux = {"c_real": channels}
            return y
        bias = None
        if use_bias:
            bias = get_variable("bias", [channels], initializer=constant_initializer(0.0))
            if c_pad != channels:
                bias = _padded_bias(bias, c_pad)
        desc.Co = c_pad          # extra output channels are exactly zero (zero weights, zero bias)
        y = _conv_op(x, wn, bias, desc, act=_act, residual=_residual, pre_relu=_pre_relu)
        if c_pad != channels:
            y.aux = {"c_real": channels}
        return y


def deconv(x, channels, opt, kernel=4, stride=2, padding='SAME', use_bias=True, scope='deconv_0',
           _residual=None):
    """ops.py:116-139 (conv2d_transpose, SAME)."""
    with variable_scope(scope):
        B, H, W, Cin = x.shape
        if padding != 'SAME':
            raise ValueError("deconv padding %s is not supported" % padding)
        if not opt.get("conv", {}).get("sn", True):
            raise ValueError("deconv with sn=False (tf.layers.conv2d_transpose path) is not supported")
        Ho, Wo = H * stride, W * stride
        tot = max((H - 1) * stride + kernel - Ho, 0)       # SAME pad of the adjoint conv
        p = tot // 2
        desc = ConvDesc(B, H, W, Cin, Ho, Wo, channels, kernel, kernel, stride, p, p, tot - p, tot - p,
                        _lib.PAD_ZERO, 1)
        w = get_variable("kernel", shape=[kernel, kernel, channels, Cin], initializer=weight_init,
                         regularizer=opt.get("conv", {}).get("regularizer", None))                                # ops.py:127
        wn = _sn_weight(w, kernel * kernel, False, x.dtype)
        bias = None
        if use_bias:
            bias = get_variable("bias", [channels], initializer=constant_initializer(0.0))
        return _conv_op(x, wn, bias, desc, residual=_residual)


def get_variable_with_custom_lr(name, shape, regularizer, lrmul):
    """ops.py:141-146 (lrmul == 1 on the hot path)."""
    w = get_variable(name, shape, initializer=truncated_normal_initializer(mean=0.0, stddev=0.02 / lrmul), regularizer=regularizer)
    if lrmul == 1.0:
        return w
    return scale(w, lrmul)


def flatten(x):
    return reshape(x, [x.shape[0], -1])                                          # ops.py:177-178


def hw_flatten(x):
    return reshape(x, [x.shape[0], -1, x.shape[-1]])                             # ops.py:180-181


def fully_connected(x, units, opt, use_bias=True, lrmul=1.0, scope='fully_0'):
    """ops.py:148-175.  fp32 in parity mode.  In bf16 tensor-core mode the spectrally normalised projections whose shapes
    fit the tcgen05 kernels (in-features % 8 == 0, units % 16 == 0: the conditional-BatchNorm gamma/beta projections
    [1040 -> C] and dense2) run as 1x1 convs over the batch rows, bf16 operands / fp32 accumulate, result cast back to
    fp32; everything else (z -> dense1, the logits) stays fp32."""
    with variable_scope(scope):
        x = flatten(x)
        Bn, channels = x.shape
        desc = ConvDesc(Bn, 1, 1, channels, 1, 1, units, 1, 1, 1, 0, 0, 0, 0, _lib.PAD_ZERO, 0)
        bias = None
        sn = opt.get("conv", {}).get("sn", True)
        tc = (sn and _state["compute_dtype"] == BF16 and not _state["force_simt"] and _lib.lib.bg_has_tensor_core_path() == 1
              and channels % 8 == 0 and channels >= 16 and units % 16 == 0)
        x = cast(x, BF16 if tc else F32)
        if sn:
            reg = opt.get("fc_regularizer", None) if 'generator' in current_scope_name() else None             # ops.py:155-158
            w = get_variable_with_custom_lr("kernel", shape=[channels, units], regularizer=reg, lrmul=lrmul)
            if use_bias:
                bias = get_variable("bias", [units], initializer=constant_initializer(0.0))
            wn = _sn_weight(w, 1, True, BF16) if tc else spectral_norm(w, _out_dtype=F32)
        else:
            with variable_scope("dense"):                                      # tf.layers.dense (:167-173)
                reg = opt.get("fc_regularizer", None) if 'generator' in current_scope_name() else None
                wn = get_variable("kernel", shape=[channels, units], initializer=weight_init, regularizer=reg)
                if use_bias:
                    bias = get_variable("bias", [units], initializer=constant_initializer(0.0))
        x4 = reshape(x, [Bn, 1, 1, channels])
        y = _conv_op(x4, wn, bias, desc)
        return cast(reshape(y, [Bn, units]), F32)


##################################################################################
# Residual-block, Self-Attention-block
##################################################################################

def resblock(x_init, channels, opt, use_bias=True, scope='resblock'):
    """ops.py:187-198."""
    with variable_scope(scope):
        with variable_scope('res1'):
            fuse = opt["act"] is relu and not opt["bn_in_d"]
            x = conv(x_init, channels, kernel=3, stride=1, pad=1, use_bias=use_bias, opt=opt,
                     _act=_lib.ACT_RELU if fuse else _lib.ACT_NONE)
            if opt["bn_in_d"]:
                x = bn(x, opt=opt)
            if not fuse:
                x = opt["act"](x)
        with variable_scope('res2'):
            if opt["bn_in_d"]:
                x = conv(x, channels, kernel=3, stride=1, pad=1, use_bias=use_bias, opt=opt)
                x = bn(x, opt=opt)
                return x + x_init
            return conv(x, channels, kernel=3, stride=1, pad=1, use_bias=use_bias, opt=opt, _residual=x_init)


def upconv(x, channels, opt, use_bias=True, _residual=None):
    """ops.py:200-218."""
    m = opt["upsampling_method"]
    if m == 'deconv3':
        return deconv(x, channels, kernel=3, stride=2, use_bias=use_bias, opt=opt, _residual=_residual)
    elif m == 'deconv4':
        return deconv(x, channels, kernel=4, stride=2, use_bias=use_bias, opt=opt, _residual=_residual)
    elif m == 'deconv6':
        return deconv(x, channels, kernel=6, stride=2, use_bias=use_bias, opt=opt, _residual=_residual)
    elif m == 'subpixel2':
        return _plus(subpixel_conv(x, channels, kernel=2, scale=2, use_bias=use_bias, opt=opt), _residual)
    elif m == 'subpixel3':
        return _plus(subpixel_conv(x, channels, kernel=3, scale=2, use_bias=use_bias, opt=opt), _residual)
    elif m == 'resize_conv':
        x = up_sample(x, 2)
        return conv(x, channels, kernel=3, stride=1, pad=1, use_bias=use_bias, opt=opt, _residual=_residual)
    elif m == 'nn':
        return _plus(up_sample(x, 2), _residual)
    else:
        raise ValueError("Invalid upsampling method specified: " + str(opt["upsampling_method"]))


def g_conv(x, channels, opt, use_bias=True, _residual=None):
    """ops.py:220-230."""
    m = opt["g_conv"]
    if m == 'deconv3':
        return deconv(x, channels, kernel=3, stride=1, use_bias=use_bias, opt=opt, _residual=_residual)
    elif m == 'deconv4':
        return deconv(x, channels, kernel=4, stride=1, use_bias=use_bias, opt=opt, _residual=_residual)
    elif m == 'conv3':
        return conv(x, channels, kernel=3, stride=1, pad=1, use_bias=use_bias, opt=opt, _residual=_residual)
    elif m == 'conv5':
        return conv(x, channels, kernel=5, stride=1, pad=2, use_bias=use_bias, opt=opt, _residual=_residual)
    else:
        raise ValueError("Invalid generator convolution type specified: " + str(opt["g_conv"]))


def resblock_up_condition(x_init, z, channels, opt, use_bias=True, scope='resblock_up'):
    """ops.py:250-266.  cond_bn + act are one fused kernel; the residual add rides in the epilogue
    of the res2 conv (the skip branch is computed first; the result is the same sum)."""
    fuse = opt["act"] is relu
    with variable_scope(scope):
        with variable_scope('skip'):
            skip = upconv(x_init, channels, use_bias=use_bias, opt=opt)
        with variable_scope('res1'):
            x = cond_bn(x_init, z, opt=opt, _act=_lib.ACT_RELU if fuse else _lib.ACT_NONE)
            if not fuse:
                x = opt["act"](x)
            x = upconv(x, channels, use_bias=use_bias, opt=opt)
        with variable_scope('res2'):
            x = cond_bn(x, z, opt=opt, _act=_lib.ACT_RELU if fuse else _lib.ACT_NONE)
            if not fuse:
                x = opt["act"](x)
            x = g_conv(x, channels, use_bias=use_bias, opt=opt, _residual=skip)
    return x


def downconv(x, channels, opt, use_bias=True, method=None, _residual=None):
    """ops.py:269-291."""
    if method is None:
        method = opt["downsampling_method"]
    if method == 'strided_conv3':
        return conv(x, channels, kernel=3, stride=2, pad=1, use_bias=use_bias, opt=opt, _residual=_residual)
    elif method == 'resize_conv1':
        x = conv(x, channels, kernel=1, stride=1, pad=0, use_bias=use_bias, opt=opt)
        return _plus(avg_pooling(x), _residual)
    elif method == 'resize_conv3':
        x = conv(x, channels, kernel=3, stride=1, pad=1, use_bias=use_bias, opt=opt)
        return _plus(avg_pooling(x), _residual)
    elif method == 'resize_conv35':
        raise ValueError("channel-sliced multi-kernel conv (ops.py:52-59) is out of scope")
    elif method == 'pool_only':
        return _plus(avg_pooling(x), _residual)
    elif method == 'max_pool_only':
        return _plus(max_pooling(x), _residual)
    else:
        raise ValueError("Invalid downsampling method specified: " + str(method))


def resblock_down(x_init, channels, opt, use_bias=True, scope='resblock_down'):
    """ops.py:293-313 (ReLU is applied to the raw image in the first block, :297-298)."""
    with variable_scope(scope):
        with variable_scope('skip'):
            skip = downconv(x_init, channels, use_bias=use_bias, opt=opt, method=opt["downsampling_method"])
        with variable_scope('res1'):
            if opt["bn_in_d"]:
                x = bn(x_init, opt=opt)
            else:
                x = x_init
            res_method = opt["downsampling_method"]
            if res_method != 'strided_conv3':
                res_method = 'resize_conv3'
            fuse = opt["act"] is relu and not opt["bn_in_d"] and res_method == 'strided_conv3'
            if fuse:                 # act -> conv -> act with both activations fused into the conv's padding copy / epilogue
                x = conv(x, channels, kernel=3, stride=2, pad=1, use_bias=use_bias, opt=opt, _act=_lib.ACT_RELU, _pre_relu=True)
            else:
                x = opt["act"](x)
                x = downconv(x, channels, use_bias=use_bias, opt=opt, method=res_method)
        with variable_scope('res2'):
            if opt["bn_in_d"]:
                x = bn(x, opt=opt)
            if not fuse:
                x = opt["act"](x)
            x = conv(x, channels, kernel=3, stride=1, pad=1, use_bias=use_bias, opt=opt, _residual=skip)
    return x


def matmul(a, b, transpose_b=False):
    """Batched tf.matmul on [B, M, K] x [B, K, N] (or [B, N, K] with transpose_b) (ops.py:481,:485)."""
    Bn, M, K = a.shape
    N = b.shape[1] if transpose_b else b.shape[2]
    dt = bg_dtype(a)
    y = Tensor(empty([Bn, M, N], a.dtype), requires_grad=a.requires_grad or b.requires_grad)
    if transpose_b:     # b [B, N, K]: element (k, n) at n*K + k
        b_rs, b_cs = 1, K
    else:               # b [B, K, N]
        b_rs, b_cs = N, 1
    call("bg_bgemm", a.data, b.data, y.data, Bn, M, N, K, K, 1, M * K, b_rs, b_cs, b.numel() // Bn,
         N, M * N, 1.0, 0, dt, stream(), flops=2.0 * Bn * M * N * K, tag="attn_bgemm_fwd")
    if recording(a, b):
        def bw():
            dy = y.grad
            if dy is None:
                return
            if a.requires_grad:       # da[M,K] = dy[M,N] . B^T
                dst, fin = grad_target(a)
                if transpose_b:   # B^T (n,k) = b[n,k]
                    call("bg_bgemm", dy, b.data, dst, Bn, M, K, N, N, 1, M * N, K, 1, N * K, K, M * K, 1.0, 0, dt, stream())
                else:             # B^T (n,k) = b[k,n]
                    call("bg_bgemm", dy, b.data, dst, Bn, M, K, N, N, 1, M * N, 1, N, K * N, K, M * K, 1.0, 0, dt, stream())
                if fin:
                    fin()
            if b.requires_grad:
                dst, fin = grad_target(b)
                if transpose_b:   # db[N,K] = dy^T[N,M] . a[M,K]
                    call("bg_bgemm", dy, a.data, dst, Bn, N, K, M, 1, N, M * N, K, 1, M * K, K, N * K, 1.0, 0, dt, stream())
                else:             # db[K,N] = a^T[K,M] . dy[M,N]
                    call("bg_bgemm", a.data, dy, dst, Bn, K, N, M, 1, K, M * K, N, 1, M * N, N, K * N, 1.0, 0, dt, stream())
                if fin:
                    fin()
        record(bw)
    return y


def softmax(x):
    """tf.nn.softmax over the last axis (ops.py:483)."""
    n = x.shape[-1]
    rows = x.numel() // n
    y = Tensor(empty(x.shape, x.dtype), requires_grad=x.requires_grad)
    call("bg_softmax_fwd", x.data, y.data, rows, n, bg_dtype(x), stream())
    if recording(x):
        def bw():
            if y.grad is None:
                return
            dst, fin = grad_target(x)
            call("bg_softmax_bwd", y.data, y.grad, dst, rows, n, bg_dtype(x), stream())
            if fin:
                fin()
        record(bw)
    return y


def _gamma_residual(gamma, o, x):
    """x = gamma * o + x (ops.py:463, :490) with scalar trainable gamma."""
    y = Tensor(empty(x.shape, x.dtype), requires_grad=gamma.requires_grad or o.requires_grad or x.requires_grad)
    n = x.numel()
    call("bg_axpby", o.data, x.data, gamma.data, 0.0, 1.0, y.data, n, bg_dtype(x), stream())
    if recording(gamma, o, x):
        def bw():
            dy = y.grad
            if dy is None:
                return
            if x.requires_grad:
                accumulate_grad(x, dy, own=False)
            if o.requires_grad:
                g = torch.empty_like(dy)
                call("bg_axpby", dy, None, gamma.data, 0.0, 0.0, g, n, bg_dtype(g), stream())
                accumulate_grad(o, g, own=True)
            if gamma.requires_grad:
                dst, fin = grad_target(gamma)
                ws = workspace(query("bg_dot_workspace_bytes", n))
                call("bg_dot", dy, o.data, dst, n, bg_dtype(dy), ws, ws.numel(), stream())
                if fin:
                    fin()
        record(bw)
    return y


def self_attention(x, channels, opt, scope='self_attention'):
    """ops.py:445-465 (full N x N attention; dead code in the reference, kept as API surface)."""
    with variable_scope(scope):
        use_bias = opt.get("self_attention_bias", False)
        f = conv(x, channels // 8, kernel=1, stride=1, opt=opt, scope='f_conv', use_bias=use_bias)
        g = conv(x, channels // 8, kernel=1, stride=1, opt=opt, scope='g_conv', use_bias=use_bias)
        h = conv(x, channels, kernel=1, stride=1, opt=opt, scope='h_conv', use_bias=use_bias)
        s = matmul(cast(hw_flatten(g), F32), cast(hw_flatten(f), F32), transpose_b=True)
        beta = softmax(s)
        o = matmul(beta, cast(hw_flatten(h), F32))
        gamma = get_variable("gamma", [1], initializer=constant_initializer(0.0))
        o = cast(reshape(o, x.shape), x.dtype)
        return _gamma_residual(gamma, o, x)


def _attn2_fused(q, k, v):
    """softmax(q k^T) v with the flash-style tcgen05 kernels (no [N, N/4] tensor is materialised)."""
    Bn, Nq, dq = q.shape
    Nk, dv = k.shape[1], v.shape[2]
    o = Tensor(empty([Bn, Nq, dv], BF16), requires_grad=q.requires_grad or k.requires_grad or v.requires_grad)
    lse = empty([Bn, Nq], F32)
    nbw = query("bg_attn2_fwd_workspace_bytes", Bn)       # per-image score bound: images under it skip the row-maximum pass
    wsf = workspace(nbw)
    call("bg_attn2_fwd_ws", q.data, k.data, v.data, o.data, lse, Bn, Nq, Nk, dq, dv, wsf, wsf.numel(), stream(),
         flops=2.0 * Bn * Nq * Nk * (dq + dv), tag="attn2_fwd")
    if recording(q, k, v):
        def bw():
            if o.grad is None:
                return
            gq, gk, gv = torch.empty_like(q.data), torch.empty_like(k.data), torch.empty_like(v.data)
            delta = empty([Bn, Nq], F32)
            nb = 0 if os.environ.get("BG_ATTN_BWD_TWO_KERNELS") == "1" else query("bg_attn2_bwd_workspace_bytes", Bn, Nq, Nk, dq)
            ws = workspace(nb) if nb else None
            call("bg_attn2_bwd", q.data, k.data, v.data, o.data, o.grad.contiguous(), lse, delta, gq, gk, gv, Bn, Nq, Nk, dq, dv,
                 ws, ws.numel() if nb else 0, stream(), flops=2.0 * Bn * Nq * Nk * (3 * dq + 4 * dv), tag="attn2_bwd")
            for t, g in ((q, gq), (k, gk), (v, gv)):
                if t.requires_grad:
                    accumulate_grad(t, g, own=True)
        record(bw)
    return o


def _split_qkv_pooled(y, widths):
    """y [B,H,W, wf+wg+wh] (the stacked f / g / h projection) -> (maxpool2(f), g, maxpool2(h)) as contiguous tensors: the two
    pooled thirds are read straight out of y (bg_maxpool2_slice_fwd), g is a channel-slice copy.  ONE backward entry
    assembles d loss / d y: the pooling backward writes its slice in place, g's gradient is copied into its slice."""
    B, H, W, Nt = y.shape
    wf, wg, wh = widths
    npix = B * H * W
    dt = bg_dtype(y)
    flat = y.data.view(-1)
    f = Tensor(empty([B, H // 2, W // 2, wf], y.dtype), requires_grad=y.requires_grad)
    g = Tensor(empty([B, H, W, wg], y.dtype), requires_grad=y.requires_grad)
    h = Tensor(empty([B, H // 2, W // 2, wh], y.dtype), requires_grad=y.requires_grad)
    call("bg_maxpool2_slice_fwd", flat, Nt, f.data, B, H, W, wf, dt, stream())
    call("bg_copy_channels", flat[wf:], Nt, g.data, wg, npix, wg, dt, stream())
    call("bg_maxpool2_slice_fwd", flat[wf + wg:], Nt, h.data, B, H, W, wh, dt, stream())
    if recording(y):
        def bw():
            if f.grad is None and g.grad is None and h.grad is None:
                return
            gy = torch.empty_like(y.data)
            gflat = gy.view(-1)
            for t, c0, w in ((f, 0, wf), (h, wf + wg, wh)):
                dy = t.grad.contiguous() if t.grad is not None else torch.zeros_like(t.data)
                call("bg_maxpool2_slice_bwd", flat[c0:], Nt, dy, gflat[c0:], Nt, B, H, W, w, dt, stream())
            dg = g.grad.contiguous() if g.grad is not None else torch.zeros_like(g.data)
            call("bg_copy_channels", dg, wg, gflat[wf:], Nt, npix, wg, dt, stream())
            accumulate_grad(y, gy, own=True)
        record(bw)
    return f, g, h


def _qkv_projection(x, channels, opt, use_bias):
    """The f / g / h 1x1 projections of self_attention_2 (ops.py:473-479) as ONE GEMM: the three spectrally normalised weight
    matrices (separate variables, separate `u`, exactly as conv() creates them) are stacked along the output-channel axis
    (f and g zero-padded to 16 channels for the tensor cores), x is read once, and the backward pass is one dgrad and one
    wgrad.  Returns (maxpool2(f), g, maxpool2(h)) (ops.py:473, :478) as contiguous tensors."""
    cin_x = x.shape[-1]
    c8, c2 = channels // 8, channels // 2
    parts = []                                      # (normalised weight, bias or None, real width, padded width)
    for scope, width, wpad in (('f_conv', c8, round_up(c8, 16)), ('g_conv', c8, round_up(c8, 16)), ('h_conv', c2, c2)):
        with variable_scope(scope):
            reg = opt.get("conv", {}).get("regularizer", None) if 'generator' in current_scope_name() else None
            w = get_variable("kernel", shape=[1, 1, cin_x, width], initializer=weight_init, regularizer=reg)
            wn = _sn_weight(w, 1, True, x.dtype, pad_c=wpad)
            bias = get_variable("bias", [width], initializer=constant_initializer(0.0)) if use_bias else None
        parts.append((wn, bias, width, wpad))
    Nt = sum(p[3] for p in parts)
    dev = x.data.device
    # stacked operands (plumbing: a few KB per phase): native [cin, Nt] for the dgrad, K-major [Nt, cin] for the forward
    w_nat = torch.cat([wn.data.reshape(cin_x, wp) for wn, _, _, wp in parts], dim=1).reshape(1, 1, cin_x, Nt)
    w_t = torch.cat([wn.aux["w_t"].reshape(-1).view(wp, cin_x) for wn, _, _, wp in parts], dim=0)
    wn_all = Tensor(w_nat, requires_grad=any(wn.requires_grad for wn, _, _, _ in parts))
    wn_all.aux = {"w_t": w_t}
    bias_all = None
    if use_bias:
        bias_all = Tensor(torch.zeros(Nt, dtype=F32, device=dev), requires_grad=any(b.requires_grad for _, b, _, _ in parts))
        c0 = 0
        for _, b, width, wp in parts:
            bias_all.data[c0:c0 + width].copy_(b.data)
            c0 += wp
    if recording(wn_all, bias_all):
        def bw():                                    # un-stack the gradients (runs after the conv's own backward entry)
            c1 = 0
            for wn, b, width, wp in parts:
                if wn.requires_grad and wn_all.grad is not None:
                    accumulate_grad(wn, wn_all.grad.reshape(cin_x, Nt)[:, c1:c1 + wp].contiguous(), own=True)
                if b is not None and b.requires_grad and bias_all.grad is not None:
                    accumulate_grad(b, bias_all.grad[c1:c1 + width].contiguous(), own=True)
                c1 += wp
        record(bw)
    B, H, W, _ = x.shape
    desc = ConvDesc(B, H, W, cin_x, H, W, Nt, 1, 1, 1, 0, 0, 0, 0, _lib.PAD_ZERO, 0)
    y = _conv_op(x, wn_all, bias_all, desc)
    f, g, h = _split_qkv_pooled(y, [p[3] for p in parts])
    if parts[0][3] != c8:
        f.aux = {"c_real": c8}
        g.aux = {"c_real": c8}
    return f, g, h


def self_attention_2(x, channels, opt, scope='self_attention'):
    """ops.py:467-492 (the live block: 2x2 max-pooled keys/values, C/8 q/k, C/2 v, 1x1 out conv)."""
    with variable_scope(scope):
        use_bias = opt.get("self_attention_bias", False)
        fused = (x.dtype == BF16 and not _state["force_simt"] and _lib.lib.bg_has_tensor_core_path() == 1
                 and (channels // 2) % 16 == 0 and channels // 2 <= 128 and channels // 8 <= 64)
        pad = 16 if fused else 0      # q/k channels zero-padded to a multiple of 16 for the tensor cores
        if fused and c_real(x) == x.shape[-1] and os.environ.get("BG_NO_QKV_FUSION") != "1":
            f, g, h = _qkv_projection(x, channels, opt, use_bias)          # f, h already max-pooled
        else:
            f = conv(x, channels // 8, kernel=1, stride=1, opt=opt, scope='f_conv', use_bias=use_bias, _pad_cout_to=pad)
            g = conv(x, channels // 8, kernel=1, stride=1, opt=opt, scope='g_conv', use_bias=use_bias, _pad_cout_to=pad)
            h = conv(x, channels // 2, kernel=1, stride=1, opt=opt, scope='h_conv', use_bias=use_bias)
            f = max_pooling(f)
            h = max_pooling(h)
        if fused:
            o = _attn2_fused(hw_flatten(g), hw_flatten(f), hw_flatten(h))           # [bs, N, C/2]
            o = reshape(o, [x.shape[0], x.shape[1], x.shape[2], channels // 2])
        else:
            # scores / softmax / PV in fp32 (bf16 storage of the logits would put a ~|s| * 2^-8 error
            # into the exponent)
            s = matmul(cast(hw_flatten(g), F32), cast(hw_flatten(f), F32), transpose_b=True)   # [bs, N, N/4], no 1/sqrt(d)
            beta = softmax(s)
            o = matmul(beta, cast(hw_flatten(h), F32))                      # [bs, N, C/2]
            o = cast(reshape(o, [x.shape[0], x.shape[1], x.shape[2], channels // 2]), x.dtype)
        gamma = get_variable("gamma", [1], initializer=constant_initializer(0.0))
        if fused and os.environ.get("BG_NO_GAMMA_FOLD") != "1":
            return _attn_out_gamma_folded(o, x, gamma, channels, opt, use_bias)
        o = conv(o, channels, kernel=1, stride=1, opt=opt, scope='attn_conv', use_bias=use_bias)
        return _gamma_residual(gamma, o, x)


def _attn_out_gamma_folded(o, x, gamma, channels, opt, use_bias):
    """x + gamma * attn_conv(o) (ops.py:486-490) as ONE conv: the scalar gamma is folded into the (tiny) normalised 1x1 weight
    and bias, x rides in the conv epilogue as the residual.  The three elementwise passes of `gamma * o + x`, the two of its
    backward and the full-size dot product for d gamma disappear; d gamma = <d(gamma W), W> + <d(gamma b), b> is taken on the
    C/2 x C weight instead."""
    cin = o.shape[-1]
    with variable_scope('attn_conv'):
        reg = opt.get("conv", {}).get("regularizer", None) if 'generator' in current_scope_name() else None
        w = get_variable("kernel", shape=[1, 1, cin, channels], initializer=weight_init, regularizer=reg)
        wn = _sn_weight(w, 1, True, o.dtype)
        bias = get_variable("bias", [channels], initializer=constant_initializer(0.0)) if use_bias else None
    dtw = bg_dtype(wn)
    nW = wn.numel()
    ws = Tensor(torch.empty_like(wn.data), requires_grad=wn.requires_grad or gamma.requires_grad)
    ws_t = torch.empty_like(wn.aux["w_t"])
    call("bg_axpby", wn.data, None, gamma.data, 0.0, 0.0, ws.data, nW, dtw, stream())
    call("bg_axpby", wn.aux["w_t"], None, gamma.data, 0.0, 0.0, ws_t, nW, dtw, stream())
    ws.aux = {"w_t": ws_t}
    bs = None
    if bias is not None:
        bs = Tensor(empty([channels], F32), requires_grad=bias.requires_grad or gamma.requires_grad)
        call("bg_axpby", bias.data, None, gamma.data, 0.0, 0.0, bs.data, channels, _lib.BG_F32, stream())
    if recording(ws, bs):
        def bw():                                    # runs after the conv's backward entry: un-fold gamma
            dws = ws.grad                            # fp32, d loss / d (gamma W)
            dbs = bs.grad if bs is not None else None
            if gamma.requires_grad:
                dg = torch.zeros(1, dtype=F32, device=gamma.data.device)
                for d, v, n in ((dws, wn.data, nW), (dbs, bias.data if bias is not None else None, channels)):
                    if d is None or v is None:
                        continue
                    v32 = v if v.dtype == F32 else empty(list(v.shape), F32)
                    if v.dtype != F32:
                        call("bg_cast", v, bg_dtype(v), v32, _lib.BG_F32, n, stream())
                    part = empty([1], F32)
                    wsd = workspace(query("bg_dot_workspace_bytes", n))
                    call("bg_dot", d, v32, part, n, _lib.BG_F32, wsd, wsd.numel(), stream())
                    call("bg_add", dg, part, dg, 1, _lib.BG_F32, stream())
                accumulate_grad(gamma, dg, own=True)
            if wn.requires_grad and dws is not None:
                g = torch.empty_like(dws)
                call("bg_axpby", dws, None, gamma.data, 0.0, 0.0, g, nW, _lib.BG_F32, stream())
                accumulate_grad(wn, g, own=True)
            if bias is not None and bias.requires_grad and dbs is not None:
                g = torch.empty_like(dbs)
                call("bg_axpby", dbs, None, gamma.data, 0.0, 0.0, g, channels, _lib.BG_F32, stream())
                accumulate_grad(bias, g, own=True)
        record(bw)
    B, H, W, _ = o.shape
    desc = ConvDesc(B, H, W, cin, H, W, channels, 1, 1, 1, 0, 0, 0, 0, _lib.PAD_ZERO, 0)
    return _conv_op(o, ws, bs, desc, residual=x)


##################################################################################
# Sampling
##################################################################################

def global_sum_pooling(x, _relu_first=False):
    """ops.py:503-506 (optionally fused with the preceding ReLU, BigGAN.py:669-671)."""
    B, H, W, C = x.shape
    y = Tensor(empty([B, C], x.dtype), requires_grad=x.requires_grad)
    call("bg_sumpool_fwd", x.data, y.data, B, H * W, C, int(_relu_first), bg_dtype(x), stream())
    if recording(x):
        def bw():
            if y.grad is None:
                return
            dst, fin = grad_target(x)
            call("bg_sumpool_bwd", x.data, y.grad, dst, B, H * W, C, int(_relu_first), bg_dtype(x), stream())
            if fin:
                fin()
        record(bw)
    return y


def max_pooling(x):
    """ops.py:508-510."""
    B, H, W, C = x.shape
    y = Tensor(empty([B, H // 2, W // 2, C], x.dtype), requires_grad=x.requires_grad)
    call("bg_maxpool2_fwd", x.data, y.data, B, H, W, C, bg_dtype(x), stream())
    if recording(x):
        def bw():
            if y.grad is None:
                return
            dst, fin = grad_target(x)
            call("bg_maxpool2_bwd", x.data, y.grad, dst, B, H, W, C, bg_dtype(x), stream())
            if fin:
                fin()
        record(bw)
    return y


##################################################################################
# Activation function
##################################################################################

def relu(x):
    """ops.py:529-530."""
    y = Tensor(empty(x.shape, x.dtype), requires_grad=x.requires_grad)
    y.aux = x.aux if isinstance(x.aux, dict) and "c_real" in x.aux else None
    call("bg_relu_fwd", x.data, y.data, x.numel(), bg_dtype(x), stream())
    if recording(x):
        def bw():
            if y.grad is None:
                return
            dst, fin = grad_target(x)
            call("bg_relu_bwd", x.data, y.grad, dst, x.numel(), bg_dtype(x), stream())
            if fin:
                fin()
        record(bw)
    return y


def _plus(x, residual):
    return x if residual is None else add(x, residual)


def _keep_c_real(y, x):
    y.aux = x.aux if isinstance(x.aux, dict) and "c_real" in x.aux else None
    return y


def lrelu(x, alpha=0.2):
    """ops.py:525-526 (tf.nn.leaky_relu)."""
    y = _keep_c_real(Tensor(empty(x.shape, x.dtype), requires_grad=x.requires_grad), x)
    call("bg_lrelu_fwd", x.data, y.data, x.numel(), float(alpha), bg_dtype(x), stream())
    if recording(x):
        def bw():
            if y.grad is None:
                return
            dst, fin = grad_target(x)
            call("bg_lrelu_bwd", x.data, y.grad, dst, x.numel(), float(alpha), bg_dtype(x), stream())
            if fin:
                fin()
        record(bw)
    return y


def prelu(x, scope=None, init_val=0.0):
    """ops.py:532-537: relu(x) + alpha * (x - |x|) / 2 with one trainable alpha per channel, created under
    '<scope>/alpha' ('prelu', 'prelu_1', ... within the enclosing scope when no scope is given)."""
    with variable_scope(scope, default_name="prelu"):
        C, Cx = c_real(x), x.shape[-1]
        alphas = get_variable('alpha', [C], initializer=constant_initializer(init_val))
        a_dev = alphas.data
        if Cx != C:                  # zero-padded image tensor (bf16 tensor-core mode): the pad channels are 0 and stay 0
            a_dev = torch.zeros(Cx, dtype=F32, device=x.data.device)
            a_dev[:C].copy_(alphas.data)
        npix = x.numel() // Cx
        y = _keep_c_real(Tensor(empty(x.shape, x.dtype), requires_grad=x.requires_grad or alphas.requires_grad), x)
        call("bg_prelu_fwd", x.data, a_dev, y.data, npix, Cx, bg_dtype(x), stream())
        if recording(x, alphas):
            def bw():
                if y.grad is None:
                    return
                dx, fin = grad_target(x) if x.requires_grad else (None, None)
                da = empty([Cx], F32)
                ws = workspace(query("bg_prelu_workspace_bytes", npix, Cx))
                call("bg_prelu_bwd", x.data, a_dev, y.grad, dx, da, 0, npix, Cx, bg_dtype(x), ws, ws.numel(), stream())
                if fin:
                    fin()
                if alphas.requires_grad:
                    accumulate_grad(alphas, da[:C].contiguous() if Cx != C else da, own=True)
            record(bw)
        return y


def avg_pooling(x):
    """ops.py:512-514: tf.layers.average_pooling2d(2, 2, 'SAME') (even maps)."""
    B, H, W, C = x.shape
    y = _keep_c_real(Tensor(empty([B, H // 2, W // 2, C], x.dtype), requires_grad=x.requires_grad), x)
    call("bg_avgpool2_fwd", x.data, y.data, B, H, W, C, bg_dtype(x), stream())
    if recording(x):
        def bw():
            if y.grad is None:
                return
            dst, fin = grad_target(x)
            call("bg_avgpool2_bwd", y.grad, dst, B, H, W, C, bg_dtype(x), stream())
            if fin:
                fin()
        record(bw)
    return y


def up_sample(x, scale_factor=2):
    """ops.py:516-519: tf.image.resize_nearest_neighbor to scale_factor times the size."""
    if scale_factor != 2:
        raise ValueError("up_sample: only scale_factor=2 is supported")
    B, H, W, C = x.shape
    y = _keep_c_real(Tensor(empty([B, 2 * H, 2 * W, C], x.dtype), requires_grad=x.requires_grad), x)
    call("bg_upsample2_fwd", x.data, y.data, B, H, W, C, bg_dtype(x), stream())
    if recording(x):
        def bw():
            if y.grad is None:
                return
            dst, fin = grad_target(x)
            call("bg_upsample2_bwd", y.grad, dst, B, H, W, C, bg_dtype(x), stream())
            if fin:
                fin()
        record(bw)
    return y


def depth_to_space(x, block_size):
    """tf.nn.depth_to_space (NHWC)."""
    B, H, W, C = x.shape
    r = int(block_size)
    y = Tensor(empty([B, H * r, W * r, C // (r * r)], x.dtype), requires_grad=x.requires_grad)
    call("bg_depth_to_space", x.data, y.data, B, H, W, C, r, 0, bg_dtype(x), stream())
    if recording(x):
        def bw():
            if y.grad is None:
                return
            dst, fin = grad_target(x)
            call("bg_depth_to_space", y.grad, dst, B, H, W, C, r, 1, bg_dtype(x), stream())
            if fin:
                fin()
        record(bw)
    return y


def subpixel_conv(x, channels, opt, kernel=3, scale=2, use_bias=True, scope='subpixel_conv_0'):
    """ops.py:23-27."""
    r2 = scale * scale
    x = conv(x, channels * r2, opt=opt, kernel=kernel, stride=1, use_bias=use_bias, pad=(kernel - 1) / 2.0, scope=scope)
    return depth_to_space(x, block_size=scale)


def tanh(x):
    """ops.py:539-540."""
    raise NotImplementedError("standalone tanh: use conv(..., _act=ACT_TANH); see model.generator")


##################################################################################
# Normalization function
##################################################################################

def _bn_core(x, gamma, beta, gb_bstride, run_mean, run_var, momentum, var_correction_fn, opt, act):
    """Shared (conditional) batch-norm forward/backward.  gamma/beta: fp32 Tensors [B,C] or [C]."""
    B, H, W, C = x.shape
    HW = H * W
    dt = bg_dtype(x)
    eps = 1e-05
    y = Tensor(empty(x.shape, x.dtype), requires_grad=x.requires_grad or gamma.requires_grad or beta.requires_grad)
    if opt["is_training"]:
        sums = empty([2 * C], torch.float64)
        ws = workspace(query("bg_bn_workspace_bytes", B, HW, C))
        call("bg_bn_stats", x.data, sums, B, HW, C, dt, ws, ws.numel(), stream())
        dist.all_reduce_sum(sums)                                  # cross-replica statistics
        count = float(B * HW * dist.world_size)
        mean = empty([C], F32)
        var = empty([C], F32)
        call("bg_bn_finalize", sums, count, mean, var, run_mean.data, run_var.data, float(momentum),
             float(var_correction_fn(count)), C, stream())
    else:
        mean, var = run_mean.data, run_var.data
        count = float(B * HW)
    call("bg_bn_apply", x.data, mean, var, gamma.data, beta.data, gb_bstride, y.data, act, eps, B, HW, C, dt, stream())
    if recording(x, gamma, beta):
        def bw():
            dy = y.grad
            if dy is None:
                return
            dgamma = empty([B, C], F32)
            dbeta = empty([B, C], F32)
            ws2 = workspace(query("bg_bn_workspace_bytes", B, HW, C))
            call("bg_bn_bwd_reduce", dy, x.data, mean, var, gamma.data, beta.data, gb_bstride, dgamma, dbeta, act, eps,
                 B, HW, C, dt, ws2, ws2.numel(), stream())
            if x.requires_grad:
                s = empty([2 * C], torch.float64)
                if opt["is_training"]:
                    call("bg_bn_bwd_sums", dgamma, dbeta, gamma.data, gb_bstride, s, B, C, stream())
                    dist.all_reduce_sum(s)
                else:
                    s.zero_()                                      # population statistics are constants
                dst, fin = grad_target(x)
                call("bg_bn_bwd_apply", dy, x.data, mean, var, gamma.data, beta.data, gb_bstride, s, count, dst, act,
                     eps, B, HW, C, dt, stream())
                if fin:
                    fin()
            if gb_bstride:
                if gamma.requires_grad:
                    accumulate_grad(gamma, dgamma, own=True)
                if beta.requires_grad:
                    accumulate_grad(beta, dbeta, own=True)
            else:
                for p, d in ((gamma, dgamma), (beta, dbeta)):
                    if p.requires_grad:
                        dst, fin = grad_target(p)
                        ws3 = workspace(query("bg_colsum_workspace_bytes", B, C))
                        call("bg_colsum", d, dst, B, C, _lib.BG_F32, ws3, ws3.numel(), stream())
                        if fin:
                            fin()
        record(bw)
    return y


class FcGroupItem(ctypes.Structure):
    """mirror of bg_fc_group_item"""
    _fields_ = [(n, ctypes.c_void_p) for n in ("x", "w", "bias", "y")] + [(n, ctypes.c_int) for n in ("K", "N", "ldw", "reserved")]


def precompute_cbn_projections(items, opt):
    """Hoists the gamma / beta projections of conditional batch norm (ops.py:623-627) out of the blocks: their inputs (z chunk
    + class vector) exist before the first block runs, so all of them are computed by ONE grouped launch here and their weight
    / bias gradients by ONE launch at the end of the backward pass (csrc/fc_group.cu) instead of ~18 launches per CBN layer.
    items: [(scope path of the batch-norm scope relative to the current scope, z Tensor [B, K] fp32, channels)].
    The results wait in the store's phase cache until `condition_batch_norm` asks for them under that scope name; the
    variables are created under exactly the names fully_connected would give them.  Returns False (nothing done) when the
    configuration does not fit the grouped kernels; the blocks then compute their own projections."""
    store = get_store()
    bn_type = opt.get("bn", {}).get("type", "bn")
    if not (_state["compute_dtype"] == BF16 and not _state["force_simt"] and _lib.lib.bg_has_tensor_core_path() == 1
            and opt.get("conv", {}).get("sn", True) and bn_type in ("bn", "batch_norm") and 0 < 2 * len(items) <= 32):
        return False
    for _, z, c in items:
        K = z.numel() // z.shape[0]
        if z.requires_grad or z.dtype != F32 or K % 4 != 0 or c % 4 != 0:       # 16-byte vector loads of the grouped kernels
            return False
    recs, arr = [], (FcGroupItem * (2 * len(items)))()
    B = items[0][1].shape[0]
    for path, z, c in items:
        K = z.numel() // B
        with contextlib.ExitStack() as es:
            for part in path.split("/"):
                es.enter_context(variable_scope(part))
            key = current_scope_name()
            pair = []
            for name in ("beta", "gamma"):                 # creation order of ops.py:623-627
                with variable_scope(name):
                    reg = opt.get("fc_regularizer", None) if 'generator' in current_scope_name() else None
                    w = get_variable_with_custom_lr("kernel", shape=[K, c], regularizer=reg, lrmul=1.0)
                    bias = get_variable("bias", [c], initializer=constant_initializer(0.0))
                    wn = _sn_weight(w, 1, True, BF16)
                y = Tensor(empty([B, c], F32), requires_grad=wn.requires_grad or bias.requires_grad)
                it = arr[len(recs)]
                it.x, it.w, it.bias, it.y = z.data.data_ptr(), wn.aux["w_t"].data_ptr(), bias.data.data_ptr(), y.data.data_ptr()
                it.K, it.N, it.ldw = K, c, K          # w_t is the [c][K] (K-major) copy, whatever shape its handle carries
                recs.append((z, wn, bias, y))
                pair.append(y)
            store.cbn_cache[key] = (pair[1], pair[0])      # (gamma, beta)
    call("bg_fc_group_fwd", arr, len(recs), B, _lib.BG_BF16, stream(), tag="fc_group_fwd")
    if recording(*[t for r in recs for t in (r[1], r[2])]):
        def bw():
            live, fins = [], []
            for z, wn, bias, y in recs:
                if y.grad is None:
                    continue
                dw, fin_w = grad_target(wn, dtype=F32) if wn.requires_grad else (None, None)
                db, fin_b = grad_target(bias) if bias.requires_grad else (None, None)
                live.append((z, y, dw, db))
                fins += [f for f in (fin_w, fin_b) if f]
            if not live:
                return
            arr2 = (FcGroupItem * len(live))()
            scratch = []
            for it, (z, y, dw, db) in zip(arr2, live):
                g = y.grad.contiguous()
                if dw is None:                              # bias-only gradient: the kernel still needs a destination
                    dw = empty([z.numel() // B, y.shape[1]], F32)
                scratch.append((g, dw))
                it.x, it.w, it.bias, it.y = z.data.data_ptr(), dw.data_ptr(), (db.data_ptr() if db is not None else None), g.data_ptr()
                it.K, it.N, it.ldw = z.numel() // B, y.shape[1], dw.shape[-1]
            call("bg_fc_group_bwd", arr2, len(live), B, stream(), tag="fc_group_bwd")
            for f in fins:
                f()
        record(bw)
    return True


def condition_batch_norm(x, z, opt={}, scope='batch_norm', _act=_lib.ACT_NONE):
    """ops.py:611-643."""
    with variable_scope(scope):
        c = x.shape[-1]
        decay = opt.get("bn", {}).get("momentum", 0.98)
        test_mean = get_variable("pop_mean", shape=[c], initializer=constant_initializer(0.0), trainable=False)
        test_var = get_variable("pop_var", shape=[c], initializer=constant_initializer(1.0), trainable=False)
        hoisted = get_store().cbn_cache.pop(current_scope_name(), None)
        if hoisted is not None:                  # projections computed up front by precompute_cbn_projections
            gamma, beta = hoisted
            return _bn_core(x, gamma, beta, c, test_mean, test_var, decay, lambda n: 1.0, opt, _act)
        if (_state["compute_dtype"] == BF16 and not _state["force_simt"] and _lib.lib.bg_has_tensor_core_path() == 1
                and z.numel() // z.shape[0] % 8 == 0 and c % 16 == 0 and opt.get("conv", {}).get("sn", True)):
            z = cast(flatten(z), BF16)           # both projections take the tensor-core path: cast the condition once
        beta = fully_connected(z, units=c, scope='beta', opt=opt)
        gamma = fully_connected(z, units=c, scope='gamma', opt=opt)
        return _bn_core(x, gamma, beta, c, test_mean, test_var, decay, lambda n: 1.0, opt, _act)


def normalize_renorm_clipping_params(renorm_clipping):
    """ops.py:587-597."""
    if "rmax" not in renorm_clipping:
        renorm_clipping["rmax"] = 1.5
    if "dmax" not in renorm_clipping:
        renorm_clipping["dmax"] = 0.5
    if "rmax" in renorm_clipping and "rmin" not in renorm_clipping:
        renorm_clipping["rmin"] = 1.0 / renorm_clipping["rmax"]
    return renorm_clipping


def _renorm_affine(gamma, beta, r, d):
    """gamma' = r gamma, beta' = beta + d gamma over [B, C] (r, d [C] are constants in the backward, ops.py:692-693)."""
    B, C = gamma.shape
    g2 = Tensor(empty([B, C], F32), requires_grad=gamma.requires_grad)
    b2 = Tensor(empty([B, C], F32), requires_grad=gamma.requires_grad or beta.requires_grad)
    call("bg_renorm_affine_fwd", gamma.data, beta.data, r, d, g2.data, b2.data, B, C, stream())
    if recording(gamma, beta):
        def bw():
            if g2.grad is None and b2.grad is None:
                return
            dg2 = g2.grad if g2.grad is not None else torch.zeros_like(gamma.data)
            db2 = b2.grad if b2.grad is not None else torch.zeros_like(beta.data)
            if gamma.requires_grad:
                dg = empty([B, C], F32)
                call("bg_renorm_affine_bwd", dg2, db2, r, d, dg, B, C, stream())
                accumulate_grad(gamma, dg, own=True)
            if beta.requires_grad:
                accumulate_grad(beta, db2, own=False)
        record(bw)
    return g2, b2


def condition_batch_renorm(x, z, opt={}, scope='batch_renorm', _act=_lib.ACT_NONE):
    """ops.py:645-715.  The correction factors r, d come from one extra statistics pass; the normalisation itself is the
    conditional batch norm kernel set with gamma' = r gamma, beta' = beta + d gamma."""
    with variable_scope(scope):
        B, H, W, c = x.shape
        renorm_clipping = normalize_renorm_clipping_params(dict(opt.get("bn", {}).get("renorm_clipping", {})))
        test_decay = opt.get("bn", {}).get("momentum", 0.98)
        renorm_decay = opt.get("bn", {}).get("renorm_momentum", 0.9)
        shared = opt.get("bn", {}).get("shared_renorm", False)
        renorm_fadein_decay = opt.get("bn", {}).get("renorm_fadein_decay", 0.9999)
        if not shared:
            test_decay = renorm_decay                                           # :658-659
        test_mean = get_variable("pop_mean", shape=[c], initializer=constant_initializer(0.0), trainable=False)
        test_var = get_variable("pop_var", shape=[c], initializer=constant_initializer(1.0), trainable=False)
        if not shared:
            renorm_mean = get_variable("renorm_mean", shape=[c], initializer=constant_initializer(0.0), trainable=False)
            renorm_var = get_variable("renorm_var", shape=[c], initializer=constant_initializer(1.0), trainable=False)
            renorm_weight = get_variable("renorm_weight", shape=[], initializer=constant_initializer(0.0), trainable=False)
        beta = fully_connected(z, units=c, scope='beta', opt=opt)
        gamma = fully_connected(z, units=c, scope='gamma', opt=opt)
        if opt["is_training"]:
            HW = H * W
            sums = empty([2 * c], torch.float64)
            ws = workspace(query("bg_bn_workspace_bytes", B, HW, c))
            call("bg_bn_stats", x.data, sums, B, HW, c, bg_dtype(x), ws, ws.numel(), stream())
            dist.all_reduce_sum(sums)
            mean, var = empty([c], F32), empty([c], F32)
            scratch_m, scratch_v = torch.zeros_like(mean), torch.ones_like(var)        # finalize also moves running statistics: discard
            call("bg_bn_finalize", sums, float(B * HW * dist.world_size), mean, var, scratch_m, scratch_v, 0.0, 1.0, c, stream())
            r, d = empty([c], F32), empty([c], F32)
            if shared:       # r, d against the population statistics as they were BEFORE this step's update (:671-673)
                call("bg_renorm_factors", mean, var, test_mean.data, test_var.data, None, None, r, d, c, 1e-05,
                     float(renorm_clipping["rmin"]), float(renorm_clipping["rmax"]), float(renorm_clipping["dmax"]), 0.0, 0.0, 0.0, 0.0, 0, stream())
            else:
                w_out = empty([1], F32)
                call("bg_renorm_factors", mean, var, renorm_mean.data, renorm_var.data, renorm_weight.data, w_out, r, d, c, 1e-05,
                     float(renorm_clipping["rmin"]), float(renorm_clipping["rmax"]), float(renorm_clipping["dmax"]),
                     float(renorm_decay), float(1 - renorm_decay), float(renorm_fadein_decay), float(1 - renorm_fadein_decay), 1, stream())
                renorm_weight.data.copy_(w_out.reshape(renorm_weight.data.shape))
            gamma, beta = _renorm_affine(gamma, beta, r, d)
        return _bn_core(x, gamma, beta, c, test_mean, test_var, test_decay, lambda n: 1.0, opt, _act)


def batch_norm(x, opt={}, scope='batch_norm', _act=_lib.ACT_NONE):
    """ops.py:580-585 (tf.layers.batch_normalization; moving_variance takes the unbiased variance)."""
    with variable_scope(scope):
        c, cx = c_real(x), x.shape[-1]               # cx > c: zero-padded image tensor of the bf16 tensor-core mode
        momentum = opt.get("bn", {}).get("momentum", 0.98)
        gamma = get_variable("gamma", [c], initializer=constant_initializer(1.0))
        beta = get_variable("beta", [c], initializer=constant_initializer(0.0))
        mm = get_variable("moving_mean", [c], initializer=constant_initializer(0.0), trainable=False)
        mv = get_variable("moving_variance", [c], initializer=constant_initializer(1.0), trainable=False)
        vec = 4 if x.dtype == F32 else 8              # the BN kernels move 16-byte channel vectors (norm.cu)
        if cx == c and c % vec == 0:
            return _bn_core(x, gamma, beta, 0, mm, mv, momentum, lambda n: n / max(n - 1.0, 1.0), opt, _act)
        # narrow tensors (bn_in_d on the 3-channel image, ops.py:295): run on zero-padded channels.  A zero channel has
        # mean 0 / variance 0 and normalises to its beta (0); its parameters and statistics are dropped again.
        cp = cx if cx != c else round_up(c, vec)
        xp = x if cx != c else pad_channels(x, cp)
        gamma_p, beta_p = _padded_param(gamma, cp, 1.0), _padded_param(beta, cp, 0.0)
        mm_p = Tensor(torch.zeros(cp, dtype=F32, device=x.data.device))
        mv_p = Tensor(torch.ones(cp, dtype=F32, device=x.data.device))
        mm_p.data[:c].copy_(mm.data)
        mv_p.data[:c].copy_(mv.data)
        y = _bn_core(xp, gamma_p, beta_p, 0, mm_p, mv_p, momentum, lambda n: n / max(n - 1.0, 1.0), opt, _act)
        mm.data.copy_(mm_p.data[:c])
        mv.data.copy_(mv_p.data[:c])
        if cx != c:                                   # stays in the padded image layout (pad channels: beta = 0)
            y.aux = {"c_real": c}
            return y
        y = slice_channels(y, c)
        y.aux = None
        return y


def bn(x, opt={}, scope='batch_norm', _act=_lib.ACT_NONE):
    """ops.py:546-561."""
    type = opt.get("bn", {}).get("type", "bn")
    if type == 'batch_norm_broken_renorm':
        type = 'batch_norm'
        if scope == 'batch_norm':
            scope = 'batch_renorm'
    if type == 'bn' or type == 'batch_norm':
        return batch_norm(x, opt=opt, scope=scope, _act=_act)
    elif type == 'batch_renorm':
        raise ValueError("BN type batch_renorm is not supported (out of scope, SURVEY 8f rank 3)")
    else:
        raise ValueError("Unknown BN type: " + str(type))


def cond_bn(x, z, opt={}, scope='batch_norm', _act=_lib.ACT_NONE):
    """ops.py:563-578."""
    type = opt.get("bn", {}).get("type", "bn")
    if type == 'batch_norm_broken_renorm':
        type = 'batch_norm'
        if scope == 'batch_norm':
            scope = 'batch_renorm'
    if type == 'bn' or type == 'batch_norm':
        return condition_batch_norm(x, z, opt=opt, scope=scope, _act=_act)
    elif type == 'batch_renorm':
        if scope == 'batch_norm':
            scope = 'batch_renorm'
        return condition_batch_renorm(x, z, opt=opt, scope=scope, _act=_act)
    else:
        raise ValueError("Unknown BN type: " + str(type))


##################################################################################
# Loss function
##################################################################################

def _seed_scaled(t, g, upstream):
    """t.grad += upstream[0] * g  (losses are scalar nodes; upstream is d(total)/d(loss))."""
    out = torch.empty_like(g)
    call("bg_axpby", g, None, upstream, 0.0, 0.0, out, out.numel(), bg_dtype(out), stream())
    accumulate_grad(t, out, own=True)


def flood_loss(loss, flood_level):
    """ops.py:847-848 -- folded into the hinge kernels (flood argument)."""
    raise NotImplementedError("flood is applied inside discriminator_loss / generator_loss")


def discriminator_loss(loss_func, real, fake, flood_level=0):
    """ops.py:753-797, 'hinge' (:788-790) + flood (:794-795).  Means are over the GLOBAL batch."""
    if loss_func != 'hinge':
        raise ValueError("loss '%s' is not supported (hinge only)" % loss_func)
    if flood_level and dist.world_size > 1:
        raise ValueError("flood_level with data parallelism needs the global loss first; unsupported")
    real32, fake32 = cast(real, F32), cast(fake, F32)
    n_r, n_f = real32.numel(), fake32.numel()
    assert n_r == n_f
    loss = Tensor(empty([1], F32), requires_grad=real.requires_grad or fake.requires_grad)
    dreal, dfake = torch.empty_like(real32.data), torch.empty_like(fake32.data)
    call("bg_hinge_d_fwd_bwd", real32.data, fake32.data, n_r, n_f, 1.0 / (n_r * dist.world_size), float(flood_level),
         loss.data, dreal, dfake, _lib.BG_F32, stream())
    if recording(real32, fake32):
        def bw():
            if loss.grad is None:
                return
            if real32.requires_grad:
                _seed_scaled(real32, dreal, loss.grad)
            if fake32.requires_grad:
                _seed_scaled(fake32, dfake, loss.grad)
        record(bw)
    return loss


def generator_loss(loss_func, fake, real, flood_level=0):
    """ops.py:799-840, 'hinge' (:832-833): -mean(fake); `real` unused."""
    if loss_func != 'hinge':
        raise ValueError("loss '%s' is not supported (hinge only)" % loss_func)
    if flood_level and dist.world_size > 1:
        raise ValueError("flood_level with data parallelism needs the global loss first; unsupported")
    fake32 = cast(fake, F32)
    n = fake32.numel()
    loss = Tensor(empty([1], F32), requires_grad=fake.requires_grad)
    dfake = torch.empty_like(fake32.data)
    call("bg_hinge_g_fwd_bwd", fake32.data, n, 1.0 / (n * dist.world_size), float(flood_level), loss.data, dfake,
         _lib.BG_F32, stream())
    if recording(fake32):
        def bw():
            if loss.grad is not None:
                _seed_scaled(fake32, dfake, loss.grad)
        record(bw)
    return loss


def cls_loss_fn(type, cls_weights):
    """utils.py:339-375, 'logistic' (:366-369).  cls_weights: device fp32 [n_labels] buffer or None."""
    if type != 'logistic':
        raise ValueError("Invalid label loss type: " + type)

    def loss(truth, answer):
        a32 = cast(answer, F32)
        Bn, nl = a32.shape
        out = Tensor(empty([1], F32), requires_grad=answer.requires_grad)
        dl = torch.empty_like(a32.data)
        tdata = truth.data if isinstance(truth, Tensor) else truth
        wsl = workspace(query("bg_logistic_cls_workspace_bytes", Bn, nl))
        call("bg_logistic_cls_fwd_bwd", a32.data, tdata, cls_weights, Bn, nl, 1.0,
             1.0 / (Bn * nl * dist.world_size), out.data, dl, 0, _lib.BG_F32, wsl, wsl.numel(), stream())
        if recording(a32):
            def bw():
                if out.grad is not None:
                    _seed_scaled(a32, dl, out.grad)
            record(bw)
        return out
    return loss
