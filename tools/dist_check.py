#!/usr/bin/env python
"""N-GPU vs 1-GPU equivalence of the data-parallel training step: run under
  python -m torch.distributed.run --nproc-per-node N --master-addr 127.0.0.1 tools/dist_check.py
Rank 0 first runs the step at the global batch on one GPU, then all ranks run it sharded; gradients, losses,
SN `u` vectors and BN running statistics must agree (fp32: 1e-4; state bit-identical across ranks)."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    from biggan_tensorflow_b200.model import Config
    from biggan_tensorflow_b200.step import Trainer
    from oracle.ref_model import Config as RC, synthetic_batch
    dtype = torch.bfloat16 if (len(sys.argv) > 1 and sys.argv[1] == "bf16") else torch.float32
    B = 8 * world
    kw = dict(img_size=32, ch=32, batch_size=B, n_labels=10, sa_gamma_init=0.1)
    sb = synthetic_batch(RC(**kw), B=B)
    dev = lambda a: torch.from_numpy(np.ascontiguousarray(a)).cuda()
    ref = None
    if rank == 0:
        t1 = Trainer(Config(**kw), compute_dtype=dtype, rank=0, world_size=1)
        dl, _ = t1.d_update(dev(sb["real"]), dev(sb["labels"]), dev(sb["z_d"]), dev(sb["labels"]))
        gd = t1.grads("discriminator")
        gl, _ = t1.g_update(dev(sb["z_g"]), dev(sb["labels"]))
        gg = t1.grads("generator")
        ref = dict(dl=float(dl.cpu()), gl=float(gl.cpu()), gd=gd, gg=gg, state=t1.store.export())
    dist.barrier()
    Bl = B // world
    sl = slice(rank * Bl, (rank + 1) * Bl)
    tn = Trainer(Config(**kw), compute_dtype=dtype, rank=rank, world_size=world, process_group=dist.group.WORLD)
    dl, _ = tn.d_update(dev(sb["real"][sl]), dev(sb["labels"][sl]), dev(sb["z_d"][sl]), dev(sb["labels"][sl]))
    gd = tn.grads("discriminator")
    gl, _ = tn.g_update(dev(sb["z_g"][sl]), dev(sb["labels"][sl]))
    gg = tn.grads("generator")
    state = tn.store.export()
    # replica consistency: u / BN running stats / parameters bit-identical across ranks
    for k in sorted(state):
        t = torch.from_numpy(state[k]).cuda()
        lo, hi = t.clone(), t.clone()
        dist.all_reduce(lo, op=dist.ReduceOp.MIN)
        dist.all_reduce(hi, op=dist.ReduceOp.MAX)
        assert torch.equal(lo, hi), "replicas diverged on " + k
    # losses are local sums scaled by 1/global batch: their sum over ranks is the global-batch loss
    lt = torch.tensor([float(dl.cpu()), float(gl.cpu())], device="cuda", dtype=torch.float64)
    dist.all_reduce(lt)
    if rank == 0:
        tol = 1e-4 if dtype == torch.float32 else 4e-2
        worst = 0.0
        for which, a, b in (("D", gd, ref["gd"]), ("G", gg, ref["gg"])):
            scale = max(np.abs(v).max() for v in b.values())
            for k in b:
                if k.endswith("f_conv/bias"):
                    continue
                e = np.abs(a[k] - b[k]).max() / max(np.abs(b[k]).max(), 1e-3 * scale)
                if dtype != torch.float32:
                    e = np.linalg.norm(a[k] - b[k]) / max(np.linalg.norm(b[k]), 1e-6)
                    if "f_conv/kernel" in k or "h_conv/kernel" in k:
                        e *= 0.25
                worst = max(worst, e)
                assert e <= tol * 3, (which, k, e)
        for k, v in ref["state"].items():
            if k.endswith("/u") or "pop_" in k or "moving_" in k:
                e = np.abs(state[k] - v).max() / max(np.abs(v).max(), 1e-6)
                assert e <= max(tol, 1e-4) * 3, (k, e)
        assert abs(lt[0].item() - ref["dl"]) <= tol * 3 * max(1, abs(ref["dl"])), (lt[0].item(), ref["dl"])
        assert abs(lt[1].item() - ref["gl"]) <= tol * 3 * max(1, abs(ref["gl"])), (lt[1].item(), ref["gl"])
        print("DIST_CHECK_OK world=%d dtype=%s worst_grad_err=%.3e d_loss=%.6f (1-GPU %.6f) g_loss=%.6f (1-GPU %.6f)"
              % (world, dtype, worst, lt[0].item(), ref["dl"], lt[1].item(), ref["gl"]))
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
