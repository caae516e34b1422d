"""DiffAugment (DiffAugment_tf.py:8-73): two-implementation agreement of the CPU oracles (runs
without a GPU) and bit-exact index parity of the CUDA kernels (gpu)."""
import numpy as np
import pytest
import torch

from oracle import ref_diffaug_np as N
from oracle import ref_ops as R


def _img(B, S, seed=0):
    return np.random.default_rng(seed).uniform(-1, 1, (B, S, S, 3)).astype(np.float32)


@pytest.mark.parametrize("S", [8, 9, 32])
def test_oracle_translation_matches_literal_numpy_exhaustively(S):
    sh = R.translation_shift(S)
    offs = [(a, b) for a in range(-sh, sh + 1) for b in range(-sh, sh + 1)]
    x = _img(len(offs), S)
    tx = np.array([o[0] for o in offs], np.int32)
    ty = np.array([o[1] for o in offs], np.int32)
    got = R.rand_translation(torch.tensor(x), {"translation_x": torch.tensor(tx), "translation_y": torch.tensor(ty)}).numpy()
    assert np.array_equal(got, N.rand_translation_np(x, tx, ty))


@pytest.mark.parametrize("S", [8, 9, 32])
def test_oracle_cutout_matches_literal_numpy_exhaustively(S):
    cs = R.cutout_size(S)
    mx = S + (1 - cs % 2)
    offs = [(a, b) for a in range(mx) for b in range(mx)]
    x = _img(len(offs), S, 1)
    ox = np.array([o[0] for o in offs], np.int32)
    oy = np.array([o[1] for o in offs], np.int32)
    got = R.rand_cutout(torch.tensor(x), {"cutout_x": torch.tensor(ox), "cutout_y": torch.tensor(oy)}).numpy()
    assert np.array_equal(got, N.rand_cutout_np(x, ox, oy))


def test_oracle_known_answers():
    x = _img(3, 16, 2)
    z = np.zeros(3, np.int32)
    t = R.rand_translation(torch.tensor(x), {"translation_x": torch.tensor(z), "translation_y": torch.tensor(z)}).numpy()
    assert np.array_equal(t, x)                                   # (0,0) translation = identity
    t = N.rand_translation_np(x, z + 2, z - 1)
    assert (t[:, -2:, :, :] == 0).all() and (t[:, :, :1, :] == 0).all()        # zero fill where shifted out
    assert np.array_equal(t[:, :-2, 1:, :], x[:, 2:, :-1, :])
    c = N.rand_cutout_np(np.ones_like(x), z + 8, z + 8)
    assert int((c[0, :, :, 0] == 0).sum()) == 8 * 8               # interior cutout zeroes cs*cs pixels
    c = N.rand_cutout_np(np.ones_like(x), z, z)
    assert int((c[0, :, :, 0] == 0).sum()) == 4 * 4               # corner: clipped rectangle area


def _rnd(B, S, seed):
    return R.draw_diffaug_randoms(np.random.Generator(np.random.PCG64(seed)), B, S, S)


@pytest.mark.gpu
@pytest.mark.parametrize("S,B", [(32, 16), (128, 8), (9, 7)])
def test_cuda_index_ops_are_bit_exact(S, B):
    """translation / cutout with the same sampled offsets: bit-exact vs the oracle and literal numpy."""
    from helpers import dev_tensor, new_product_store
    from biggan_tensorflow_b200 import DiffAugment_tf as D
    new_product_store(torch.float32)
    x = _img(B, S, 3)
    rnd = _rnd(B, S, 5)
    for policy in ("translation", "cutout", "translation,cutout"):
        D.push_randoms(rnd)
        y = D.DiffAugment(dev_tensor(x, requires_grad=False), policy=policy).numpy()
        ref = R.DiffAugment(torch.tensor(x), policy=policy, rnd=R.rnd_to_torch(rnd)).numpy()
        assert np.array_equal(y, ref), policy
    lit = N.rand_cutout_np(N.rand_translation_np(x, rnd["translation_x"], rnd["translation_y"]), rnd["cutout_x"], rnd["cutout_y"])
    assert np.array_equal(y, lit)


@pytest.mark.gpu
@pytest.mark.parametrize("policy", ["color", "color,translation,cutout", "translation,cutout", "cutout,color"])
@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
def test_cuda_diffaugment_forward_backward(policy, dtype):
    from helpers import TOL, assert_close, dev_tensor, new_product_store, seed_backward
    from biggan_tensorflow_b200 import DiffAugment_tf as D
    new_product_store(dtype)
    B, S = 6, 32
    x = _img(B, S, 4)
    rnd = _rnd(B, S, 6)
    D.push_randoms(rnd)
    tx = dev_tensor(x, dtype)
    y = D.DiffAugment(tx, policy=policy)
    dy = np.random.default_rng(7).standard_normal(x.shape).astype(np.float32)
    seed_backward(y, dy)
    xr = torch.tensor(x, dtype=torch.float64, requires_grad=True)
    yr = R.DiffAugment(xr, policy=policy, rnd=R.rnd_to_torch(rnd))
    (g,) = torch.autograd.grad((yr * torch.tensor(dy, dtype=torch.float64)).sum(), [xr])
    tol = 1e-5 if dtype == torch.float32 else TOL[dtype]
    assert_close(y.numpy(), yr.detach().numpy(), tol, "fwd")
    assert_close(tx.grad.float().cpu().numpy(), g.numpy(), tol, "bwd")
    if dtype == torch.float32:      # the zero pattern (index work) must be exact
        assert np.array_equal(y.numpy() == 0, yr.detach().numpy() == 0)


@pytest.mark.gpu
def test_cuda_random_draws_stay_in_reference_ranges():
    from biggan_tensorflow_b200 import DiffAugment_tf as D
    r = D.draw_randoms(np.random.Generator(np.random.PCG64(0)), 4096, 128, 128)
    assert r["translation_x"].min() == -16 and r["translation_x"].max() == 16
    assert r["cutout_x"].min() == 0 and r["cutout_x"].max() == 128
    assert 0 <= r["brightness"].min() and r["brightness"].max() < 1
