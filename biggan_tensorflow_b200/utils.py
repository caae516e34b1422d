"""The slice of the reference's utils.py that sits on the training step: the orthogonal
regularisers of the generator weights (utils.py:180-235; selected by `--g_regularization`, BigGAN.py:257-271, summed
into the G loss at BigGAN.py:897-898).  Same factory names and call convention as the reference: the factory returns
a callable that maps a weight variable to a scalar loss tensor.

    reg = W^T W - I                                  ('ortho')
    reg = cosine_similarity(W^T W, 1 - I)            ('ortho_cosine')
    loss = scale * l2_loss(reg) = scale * sum(reg^2) / 2,      W reshaped to [-1, last dim]

The Gram matrix and dLoss/dW = W (dG + dG^T) are fp32 GEMMs (`bg_bgemm`), the row-wise part (normalisation, loss,
dLoss/dG) is `bg_ortho_reg_rows`.
"""
import torch

from ._lib import call
from .tensor import F32, Tensor, empty, grad_target, record, recording, stream

_KINDS = {"ortho": 0, "ortho_cosine": 1}


def _gram(w2, K, C):
    """G [C, C] = W^T W for W [K, C] fp32."""
    G = empty([C, C], F32)
    # A[m, k] = W[k, m] (row stride 1, column stride C);  B[k, n] = W[k, n]
    call("bg_bgemm", w2, w2, G, 1, C, C, K, 1, C, 0, C, 1, 0, C, 0, 1.0, 0, 0, stream(), flops=2.0 * K * C * C, tag="ortho_gram")
    return G


def _add_weight_grad(w2, dG, dst, K, C, alpha, accumulate):
    """dst [K, C] (+)= alpha * W (dG + dG^T)."""
    call("bg_bgemm", w2, dG, dst, 1, K, C, C, C, 1, 0, C, 1, 0, C, 0, float(alpha), int(accumulate), 0, stream(),
         flops=2.0 * K * C * C, tag="ortho_dw")
    call("bg_bgemm", w2, dG, dst, 1, K, C, C, C, 1, 0, 1, C, 0, C, 0, float(alpha), 1, 0, stream(),
         flops=2.0 * K * C * C, tag="ortho_dw")


class OrthoRegularizer:
    def __init__(self, scale, type):
        self.scale, self.type = float(scale), type

    def _kind(self):
        if self.type not in _KINDS:
            raise ValueError("Unknown regularization method.")                    # utils.py:204, :228
        return _KINDS[self.type]

    def evaluate(self, w_data, loss_buf, dW=None, grad_scale=1.0):
        """loss_buf[0] += scale * l2_loss(reg(W)); when dW is given, dW += grad_scale * dLoss/dW.  Returns dLoss/dG."""
        kind = self._kind()
        C = w_data.shape[-1]
        K = w_data.numel() // C
        w2 = w_data.reshape(K, C)
        G = _gram(w2, K, C)
        dG = empty([C, C], F32)
        rowloss = empty([C], F32)
        call("bg_ortho_reg_rows", G, dG, rowloss, loss_buf, 1, C, self.scale, kind, stream(), tag="ortho_rows")
        if dW is not None:
            _add_weight_grad(w2, dG, dW.reshape(K, C), K, C, grad_scale, 1)
        return dG

    def __call__(self, w):
        """Reference call convention: regularizer(variable) -> scalar loss tensor (differentiable w.r.t. the variable)."""
        if w.data.dtype != F32:
            raise ValueError("regularisers act on the fp32 master weights")
        loss = Tensor(torch.zeros(1, dtype=F32, device=w.data.device), requires_grad=w.requires_grad)
        dG = self.evaluate(w.data, loss.data)
        if recording(w):
            def bw():
                if loss.grad is None:
                    return
                C = w.shape[-1]
                K = w.numel() // C
                unit = empty([K, C], F32)
                _add_weight_grad(w.data.reshape(K, C), dG, unit, K, C, 1.0, 0)
                dst, fin = grad_target(w)
                call("bg_axpby", unit, None, loss.grad, 1.0, 0.0, dst, unit.numel(), 0, stream())   # upstream scalar on the device
                if fin:
                    fin()
            record(bw)
        return loss


def orthogonal_regularizer(scale, type='ortho'):
    """utils.py:185-210 (conv kernels [kh, kw, a, c] -> [-1, c])."""
    return OrthoRegularizer(scale, type)


def orthogonal_regularizer_fc(scale, type='ortho'):
    """utils.py:212-235 (fully connected kernels [in, c])."""
    return OrthoRegularizer(scale, type)
