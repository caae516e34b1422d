"""The REAL data-parallel code of `step.Trainer` (batch sharding, losses scaled by the global batch, gradient all-reduce
over the flat arenas, cross-replica BatchNorm sums, 1/world share of the regulariser gradient, replica-consistent state)
at world_size 2 over `gloo` on CPU, fp32: each rank points `_lib.lib` at the host library (tests/host_emul) and runs one D
phase and one G phase on its half of the batch; the summed gradients, the summed losses and the state must equal the
single-process run at the global batch -- the check tools/dist_check.py makes on N GPUs over NCCL."""
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "host_emul"))
import host_env                                             # noqa: E402
from oracle import ref_model as M                           # noqa: E402


CASES = {
    "canonical": dict(img_size=32, ch=16, batch_size=4, n_labels=4, z_dim=64, sa_gamma_init=0.1),
    "cli_defaults": dict(img_size=32, ch=16, batch_size=4, n_labels=4, z_dim=64, sa_gamma_init=0.1, activation="prelu",
                         g_regularization="ortho_cosine", g_regularization_factor=1e-2),
    "bn_in_d": dict(img_size=32, ch=16, batch_size=4, n_labels=4, z_dim=64, sa_gamma_init=0.1, bn_in_d=True, activation="lrelu"),
}


def _run(kw, rank, world, pg):
    from biggan_tensorflow_b200.model import Config
    from biggan_tensorflow_b200.step import Trainer
    okw = {k: v for k, v in kw.items()}
    b = M.synthetic_batch(M.Config(**okw), B=kw["batch_size"])
    Bl = kw["batch_size"] // world
    sl = slice(rank * Bl, (rank + 1) * Bl)
    t = lambda a: torch.from_numpy(np.ascontiguousarray(a[sl]))
    tr = Trainer(Config(**kw), compute_dtype=torch.float32, rank=rank, world_size=world, process_group=pg)
    dl, _ = tr.d_update(t(b["real"]), t(b["labels"]), t(b["z_d"]), t(b["labels"]), apply=False)
    gd = tr.grads("discriminator")
    state_d = tr.store.export()
    gl, _ = tr.g_update(t(b["z_g"]), t(b["labels"]), apply=False)
    gg = tr.grads("generator")
    return dict(dl=float(dl), gl=float(gl), gd=gd, gg=gg, state_d=state_d, state_g=tr.store.export())


def _train(kw, rank, world, pg, shard, steps=2, ckpt_dir=None):
    """`steps` full iterations (D update + G update with Adam / EMA applied) -> parameters and optimizer state."""
    from biggan_tensorflow_b200.model import Config
    from biggan_tensorflow_b200.step import Trainer
    b = M.synthetic_batch(M.Config(**kw), B=kw["batch_size"])
    Bl = kw["batch_size"] // world
    sl = slice(rank * Bl, (rank + 1) * Bl)
    t = lambda a: torch.from_numpy(np.ascontiguousarray(a[sl]))
    tr = Trainer(Config(**kw), compute_dtype=torch.float32, rank=rank, world_size=world, process_group=pg, shard_optimizer=shard)
    for _ in range(steps):
        tr.train_step(t(b["real"]), t(b["labels"]), t(b["z_d"]), t(b["z_g"]), t(b["labels"]))
    out = {"vars": tr.store.export()}
    for net in ("generator", "discriminator"):
        n = sum(v.data.numel() for v in tr.store.arenas[net]["vars"])          # without the shard padding at the end
        for key in ("m", "v") + (("ema",) if net == "generator" else ()):
            out[net + "/" + key] = tr.full_opt(net, key).clone().numpy()
        out[net + "/numel"] = n
    if ckpt_dir:           # collective: every rank calls save (the slots are gathered), rank 0 alone writes, all leave by a barrier
        from biggan_tensorflow_b200 import checkpoint as ck
        d = os.path.join(ckpt_dir, "ckpt_shard%d" % int(shard))
        prefix = ck.save(tr, d, step=steps)
        assert sorted(os.listdir(d)) == ["BigGAN.model-%d.data-00000-of-00001" % steps, "BigGAN.model-%d.index" % steps, "checkpoint"]
        out["ckpt"] = ck.read_bundle(prefix)
    return out


def _worker_train(rank, world, port, so, kw, shard, out):
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    torch.set_num_threads(1)
    host_env.activate(so)
    torch.save(_train(kw, rank, world, dist.group.WORLD, shard, ckpt_dir=os.path.dirname(out)), out % rank)
    dist.barrier()
    dist.destroy_process_group()


def test_sharded_optimizer_equals_replicated_optimizer(tmp_path, host_so):
    so = host_so
    """`Trainer(shard_optimizer=True)`: reduce-scatter of the gradient arena, Adam(+EMA) on this rank's slice, all-gather of the
    parameters -- after two iterations every parameter, Adam slot and EMA shadow is BIT-IDENTICAL to the replicated optimizer
    (all-reduce + full Adam on every rank), on both ranks."""
    kw = CASES["canonical"]
    res = {}
    for shard in (False, True):
        out = str(tmp_path / ("shard%d_rank%%d.pt" % shard))
        mp.spawn(_worker_train, args=(2, 29800 + os.getpid() % 97 + shard, so, kw, shard, out), nprocs=2, join=True)
        res[shard] = [torch.load(out % r, weights_only=False) for r in (0, 1)]
    for r in (0, 1):
        a, b = res[False][r], res[True][r]
        assert sorted(a["vars"]) == sorted(b["vars"])
        for k in a["vars"]:
            assert np.array_equal(a["vars"][k], b["vars"][k]), k
        for net in ("generator", "discriminator"):
            n = a[net + "/numel"]
            for key in ("m", "v") + (("ema",) if net == "generator" else ()):
                assert np.array_equal(a[net + "/" + key][:n], b[net + "/" + key][:n]), (net, key)
        assert sorted(a["ckpt"]) == sorted(b["ckpt"])                           # the checkpoint a sharded run writes is the same file content
        for k in a["ckpt"]:
            assert np.array_equal(a["ckpt"][k], b["ckpt"][k]), k
    for k in res[True][0]["vars"]:                                              # replicas agree
        assert np.array_equal(res[True][0]["vars"][k], res[True][1]["vars"][k]), k


def _worker(rank, world, port, so, kw, out):
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    torch.set_num_threads(1)
    host_env.activate(so)
    r = _run(kw, rank, world, dist.group.WORLD)
    # losses are local sums over the GLOBAL count (the regulariser: a 1/world share per rank)
    lt = torch.tensor([r["dl"], r["gl"]], dtype=torch.float64)
    dist.all_reduce(lt)
    torch.save(dict(r, dl_sum=float(lt[0]), gl_sum=float(lt[1])), out % rank)
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("name", ["cli_defaults", "bn_in_d"])      # "canonical" runs in the sharded-optimizer test below
def test_two_rank_trainer_equals_single_process(tmp_path, host_so, name):
    so = host_so
    kw = CASES[name]
    out = str(tmp_path / "rank%d.pt")
    mp.spawn(_worker, args=(2, 29900 + os.getpid() % 97, so, kw, out), nprocs=2, join=True)
    r0, r1 = torch.load(out % 0, weights_only=False), torch.load(out % 1, weights_only=False)
    # replicas: identical gradients after the all-reduce, identical state (u, running statistics)
    for k in r0["gd"]:
        assert np.array_equal(r0["gd"][k], r1["gd"][k]), k
    for k in r0["gg"]:
        assert np.array_equal(r0["gg"][k], r1["gg"][k]), k
    for k in r0["state_g"]:
        assert np.array_equal(r0["state_g"][k], r1["state_g"][k]), k
    # single process at the global batch (same host library, this process)
    mpatch = pytest.MonkeyPatch()
    try:
        host_env.activate(so, mpatch.setattr)
        ref = _run(kw, 0, 1, None)
    finally:
        mpatch.undo()
    assert abs(r0["dl_sum"] - ref["dl"]) <= 2e-5 * max(1.0, abs(ref["dl"]))
    assert abs(r0["gl_sum"] - ref["gl"]) <= 2e-5 * max(1.0, abs(ref["gl"]))         # incl. each rank's 1/world share of the regulariser
    for which in ("gd", "gg"):
        scale = max(float(np.abs(v).max()) for v in ref[which].values())
        for k, v in ref[which].items():
            err = float(np.abs(r0[which][k] - v).max()) / max(float(np.abs(v).max()), 1e-3 * scale)
            assert err <= 3e-4, (which, k, err)
    for tag in ("state_d", "state_g"):
        for k, v in ref[tag].items():
            if k.endswith("/u") or "pop_" in k or "moving_" in k:
                assert float(np.abs(r0[tag][k] - v).max()) <= 3e-4 * max(float(np.abs(v).max()), 1e-6), (tag, k)
