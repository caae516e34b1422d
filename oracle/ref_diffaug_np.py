"""CPU ORACLE -- TEST INFRASTRUCTURE ONLY.  Independent, literal numpy restatement of the integer
index work of DiffAugment_tf.py (pad + gather_nd + scatter_nd, lines 40-66), written op-by-op the way
the TF graph executes it, so that the closed forms used by oracle/ref_ops.py and by the CUDA kernels
can be checked against a second implementation (two-implementation agreement; both also reproduce the vectors the
reference's own DiffAugment_tf.py yields over the eager TF-1 stand-in, tests/golden/ref_diffaug.npz)."""
import numpy as np


def rand_translation_np(x, tx, ty, ratio=0.125):
    """DiffAugment_tf.py:40-50, literally."""
    B, H, W, C = x.shape
    grid_x = np.clip(np.arange(H, dtype=np.int32)[None, :] + tx.reshape(B, 1) + 1, 0, H + 1)          # :46
    grid_y = np.clip(np.arange(W, dtype=np.int32)[None, :] + ty.reshape(B, 1) + 1, 0, W + 1)          # :47
    xp = np.pad(x, [[0, 0], [1, 1], [0, 0], [0, 0]])                                                  # :48
    x1 = np.stack([xp[b][grid_x[b]] for b in range(B)])                # gather_nd(batch_dims=1)
    xt = np.transpose(x1, [0, 2, 1, 3])
    xtp = np.pad(xt, [[0, 0], [1, 1], [0, 0], [0, 0]])                                                # :49
    x2 = np.stack([xtp[b][grid_y[b]] for b in range(B)])
    return np.transpose(x2, [0, 2, 1, 3])


def rand_cutout_np(x, ox, oy, ratio=0.5):
    """DiffAugment_tf.py:53-66, literally (meshgrid + clip + scatter_nd)."""
    B, H, W, C = x.shape
    cs0, cs1 = int(H * ratio + 0.5), int(W * ratio + 0.5)
    gb, gx, gy = np.meshgrid(np.arange(B), np.arange(cs0), np.arange(cs1), indexing="ij")
    gx = gx + ox.reshape(B, 1, 1) - cs0 // 2
    gy = gy + oy.reshape(B, 1, 1) - cs1 // 2
    gx = np.minimum(np.maximum(gx, 0), H - 1)
    gy = np.minimum(np.maximum(gy, 0), W - 1)
    scat = np.zeros((B, H, W), np.float32)
    np.add.at(scat, (gb.reshape(-1), gx.reshape(-1), gy.reshape(-1)), 1.0)   # scatter_nd accumulates
    mask = np.maximum(1 - scat, 0)
    return x * mask[..., None]
