"""world_size-2 `gloo` tests (CPU) of the data-parallel design of the step: batch sharding,
loss scaling by the GLOBAL batch, gradient all-reduce, and the (sum, sumsq) merge that gives cross-replica
BatchNorm its global-batch statistics.  The math under test is the oracle's (the CUDA path needs a GPU; its
N-GPU vs 1-GPU equivalence is tools/dist_check.py under `gpurun --gpus N`)."""
import os

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from oracle import ref_model as M
from oracle import ref_ops as R


def shard(rank, world, B):
    """Contiguous split of the global batch: rank r owns [r*B/N, (r+1)*B/N)."""
    Bl = B // world
    return slice(rank * Bl, (rank + 1) * Bl)


def _worker(rank, world, port, out):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    torch.set_num_threads(1)
    cfg = M.Config(img_size=32, ch=8, batch_size=4, n_labels=4, sa_gamma_init=0.1)
    b = M.synthetic_batch(cfg, B=4)
    sl = shard(rank, world, 4)
    # --- D update: no batch statistics in D -> sharded gradients, scaled by local/global batch, sum to the global gradient
    st = R.RefStore(torch.float64)
    fake_full = M.generator(R.RefStore(torch.float64), cfg, torch.tensor(b["z_d"], dtype=torch.float64),
                            torch.tensor(b["labels"], dtype=torch.float64)).detach()      # global-batch BN statistics
    real = torch.tensor(b["real"][sl], dtype=torch.float64)
    d_real = M.discriminator(st, cfg, real)
    d_fake = M.discriminator(st, cfg, fake_full[sl])
    local = R.discriminator_loss("hinge", d_real["real"], d_fake["real"]) \
        + cfg.d_cls_loss_weight * R.cls_loss_logistic(torch.tensor(b["labels"][sl], dtype=torch.float64), d_real["cls"])
    local = local * (real.shape[0] / 4.0)                   # sum over local samples / GLOBAL batch
    names = sorted(st.trainables("discriminator"))
    grads = torch.autograd.grad(local, [st.vars[n] for n in names])
    flat = torch.cat([g.reshape(-1) for g in grads])
    dist.all_reduce(flat)                                   # the gradient all-reduce of the step
    loss = local.detach().clone()
    dist.all_reduce(loss)
    # --- sync-BN statistics: all-reduce of (sum, sumsq) == moments of the global batch
    x = torch.tensor(np.random.default_rng(7).standard_normal((4, 4, 4, 6)) * 2 + 0.5)
    xs = x[sl]
    sums = torch.stack([xs.sum(dim=(0, 1, 2)), (xs * xs).sum(dim=(0, 1, 2))])
    dist.all_reduce(sums)
    n = float(4 * 4 * 4)
    mean, var = sums[0] / n, sums[1] / n - (sums[0] / n) ** 2
    if rank == 0:
        torch.save({"flat": flat, "loss": loss, "names": names, "mean": mean, "var": var}, out)
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_gloo_matches_single_process(tmp_path):
    out = str(tmp_path / "r0.pt")
    port = 29600 + os.getpid() % 300
    mp.spawn(_worker, args=(2, port, out), nprocs=2, join=True)
    got = torch.load(out)
    cfg = M.Config(img_size=32, ch=8, batch_size=4, n_labels=4, sa_gamma_init=0.1)
    b = M.synthetic_batch(cfg, B=4)
    st = R.RefStore(torch.float64)
    loss, grads, _ = M.d_update(st, cfg, b["real"], b["labels"], b["z_d"], b["labels"], apply=False)
    ref = torch.cat([grads[n].reshape(-1) for n in got["names"]])
    assert abs(float(got["loss"]) - loss) < 1e-10
    assert float((got["flat"] - ref).abs().max()) < 1e-10
    x = torch.tensor(np.random.default_rng(7).standard_normal((4, 4, 4, 6)) * 2 + 0.5)
    assert float((got["mean"] - x.mean(dim=(0, 1, 2))).abs().max()) < 1e-12
    assert float((got["var"] - x.var(dim=(0, 1, 2), unbiased=False)).abs().max()) < 1e-12


def test_shard_ranges_cover_the_batch():
    for world in (1, 2, 4, 8):
        idx = np.concatenate([np.arange(256)[shard(r, world, 256)] for r in range(world)])
        assert np.array_equal(idx, np.arange(256))
