"""One BigGAN training iteration = D update then G update (BigGAN.train, BigGAN.py:1061-1084), with the
loss / optimizer wiring of BigGAN.build_model (BigGAN.py:806-930) and utils.create_train_ops
(utils.py:252-267): hinge losses (+ logistic class loss when n_labels > 0), Adam for D, Adam +
MovingAverageOptimizer (EMA shadows) for G.

Data parallel (new design): one process per GPU, the global batch is split contiguously
across ranks; gradients are summed with one NCCL all-reduce per flat arena, BN statistics with
small all-reduces inside ops._bn_core; losses are scaled by 1/global batch so summed gradients equal
the single-device gradient of the global-batch mean.
"""
import math
import os

import torch

from . import DiffAugment_tf, _lib, ops
from ._lib import call
from .model import Config, discriminator, generator
from .tensor import (BF16, F32, Tensor, accumulate_grad, backward, bg_dtype, clear_tape, device, empty, no_grad, record,
                     recording, stream)
from .tf_shim import VariableStore, set_store


def split_rows(t, sizes):
    """Split [sum(sizes), ...] along axis 0 into views; backward reassembles the gradient."""
    outs, o = [], 0
    for s in sizes:
        outs.append(Tensor(t.data[o:o + s], requires_grad=t.requires_grad))
        o += s
    if recording(t):
        def bw():
            if all(c.grad is None for c in outs):
                return
            g = torch.empty_like(t.data)
            o2 = 0
            for c, s in zip(outs, sizes):
                if c.grad is None:
                    g[o2:o2 + s].zero_()
                else:
                    g[o2:o2 + s].copy_(c.grad)
                o2 += s
            accumulate_grad(t, g, own=True)
        record(bw)
    return outs


class Trainer:
    def __init__(self, cfg: Config, compute_dtype=F32, rank=0, world_size=1, process_group=None, shard_optimizer=None, overlap=True):
        """shard_optimizer (default: on when world_size > 1; SURVEY 8e "optionally optimizer-sharded"): the update phases
        reduce-scatter the gradient arena, run Adam(+EMA) on this rank's 1/world slice and all-gather the parameters --
        the same bytes on the wire as the all-reduce, 1/world of the optimizer's HBM traffic per rank.
        overlap (world_size > 1, CUDA): the D update's collective + Adam run on a side stream underneath the G phase's
        generator forward (which does not touch D); the BN-statistics all-reduces then use their own communicator so
        that they do not queue behind the gradient collective."""
        self.cfg = cfg
        self.compute_dtype = compute_dtype
        self.rank, self.world = rank, world_size
        self.pg = process_group
        self.pg_small = process_group
        self.shard = (world_size > 1) if shard_optimizer is None else (bool(shard_optimizer) and world_size > 1)
        self.overlap = bool(overlap) and world_size > 1
        self.comm_stream = None          # created in build() on a CUDA device
        self.peer = None                 # comm.PeerComm (NVLink peer-memory all-reduce of the BN statistics)
        self._pending = False            # D-update collective / Adam in flight on the side stream
        self.store = set_store(VariableStore())
        ops.set_compute_dtype(compute_dtype)
        assert cfg.batch_size % world_size == 0, "global batch must divide across ranks"
        self.local_batch = cfg.batch_size // world_size
        ops.dist.world_size = world_size
        if world_size > 1:
            import torch.distributed as dist

            def allreduce(buf):
                with _lib.timed("nccl_allreduce_bn_stats"):
                    dist.all_reduce(buf, group=self.pg_small)
                return buf
            ops.dist.all_reduce_sum = allreduce
        else:
            ops.dist.all_reduce_sum = lambda b: b
        import numpy as _np
        DiffAugment_tf.random_source = _np.random.Generator(_np.random.PCG64(4 + rank))   # per-rank augmentation draws
        self.t_d = 0
        self.t_g = 0
        self.opt = {}
        self.cls_weights = None
        self._built = False

    # ------------------------------------------------------------------------------------------
    def build(self):
        """Create every variable (dry forward on one sample), then pack each network's trainables into
        flat arenas (params / grads / Adam m, v / EMA shadows)."""
        cfg = self.cfg
        dev = device()
        if self.overlap and dev.type == "cuda" and self.comm_stream is None:
            import torch.distributed as dist
            self.comm_stream = torch.cuda.Stream()
            self.pg_small = dist.new_group(ranks=list(range(self.world)))      # own communicator for the BN statistics
        if self.world > 1 and dev.type == "cuda" and self.peer is None and os.environ.get("BG_BN_SYNC", "p2p") == "p2p":
            # cross-replica BN statistics: one-shot all-reduce over NVLink peer memory (csrc/comm.cu) instead of NCCL
            from .comm import MAX_DOUBLES, PeerComm
            peer = PeerComm(self.rank, self.world, self.pg)          # collective; never raises on a local failure
            if peer.error is not None:
                print("[biggan_b200] rank %d: peer-memory exchange unavailable (%r)" % (self.rank, peer.error))
            if peer.agree():                                         # every rank takes the same decision
                self.peer = peer
                nccl_sum = ops.dist.all_reduce_sum

                def peer_sum(buf):
                    if buf.dtype == torch.float64 and buf.numel() <= MAX_DOUBLES:
                        return self.peer.all_reduce_sum(buf)
                    return nccl_sum(buf)
                ops.dist.all_reduce_sum = peer_sum
            else:                                                    # BN statistics stay on NCCL
                peer.close()
        z = torch.zeros(2, 1, 1, cfg.z_dim, device=dev)
        cls = torch.zeros(2, max(cfg.n_labels, 1), device=dev)
        set_store(self.store)
        ops.set_compute_dtype(self.compute_dtype)
        self.store.begin_phase()
        self.store.sn_registry = {}                  # enumerate every spectral_norm call of both networks
        with no_grad():
            fake = generator(cfg, z, cls)
            discriminator(cfg, fake)
        registry, self.store.sn_registry = self.store.sn_registry, None
        snap = {k: v.data.clone() for k, v in self.store.vars.items() if not v.trainable}
        for k, d in snap.items():       # undo the dry run's side effects on u / BN running stats
            self.store.vars[k].data.copy_(torch.from_numpy(self._initial(k, d.shape)).to(dev))
        for net in ("generator", "discriminator"):
            ar = self.store.pack(net, multiple=4 * self.world if self.shard else 1)
            n = ar["params"].numel()
            self.opt[net] = {"m": torch.zeros(n, device=dev), "v": torch.zeros(n, device=dev),
                             "ema": ar["params"].clone() if net == "generator" else None,
                             "state": torch.zeros(2, device=dev),          # [t, lr_t]
                             # sharded optimizer: persistent landing buffers of the reduce-scatter / source of the all-gather
                             "gshard": torch.zeros(n // self.world, device=dev) if self.shard else None,
                             "pshard": torch.zeros(n // self.world, device=dev) if self.shard else None}
        from .sn_group import SnGroup
        self.sn_groups = {net: SnGroup([e for e in registry.values() if net in e["u"].name])
                          for net in ("generator", "discriminator")}
        if cfg.acgan:
            self.cls_weights = torch.ones(cfg.n_labels, device=dev)            # BigGAN.py:100
        # one regularisation loss per generator weight created with a regulariser (BigGAN.py:257-271, :897-898)
        self.regs = [(self.store.vars[n], r) for n, r in self.store.regularizers.items()]
        self.reg_loss = None
        self.store.begin_phase()
        clear_tape()
        self._built = True

    def _initial(self, name, shape):
        from .tf_shim import constant_initializer, random_normal_initializer
        if name.endswith("/u"):
            return random_normal_initializer()(name, shape)
        if name.endswith("pop_var") or name.endswith("moving_variance"):
            return constant_initializer(1.0)(name, shape)
        return constant_initializer(0.0)(name, shape)

    def load(self, params):
        self.store.load(params)
        if self.opt.get("generator") and self.opt["generator"]["ema"] is not None:
            self.opt["generator"]["ema"].copy_(self.store.arenas["generator"]["params"])

    # ------------------------------------------------------------------------------------------
    def _set_phase(self, net):
        set_store(self.store)                        # the variable store is process-global (like TF's default graph)
        for k, v in self.store.vars.items():
            v.requires_grad = bool(v.trainable and net in k)
            v.grad = None
            v.grad_set = False
            v.grad_owned = True
            v.n_contribs = v.n_masked = 0
        self.store.begin_phase()
        clear_tape()
        # every forward of a network advances its u once per phase.  In the G phase D's spectral norm waits until the
        # generator forward is issued (`_sn_forward_discriminator`): D's parameters may still be in the D update's
        # collective / Adam on the side stream
        self.sn_groups["generator"].forward(self.store)
        if net == "discriminator":
            self._sn_forward_discriminator()

    def _sn_forward_discriminator(self):
        self._join_side()
        self.sn_groups["discriminator"].forward(self.store)

    def _join_side(self):
        """Main stream waits for the side stream's D update (no-op when nothing is in flight)."""
        if self._pending:
            torch.cuda.current_stream().wait_stream(self.comm_stream)
            self._pending = False

    def _shard_slice(self, net):
        n = self.store.arenas[net]["params"].numel() // self.world
        return slice(self.rank * n, (self.rank + 1) * n)

    def _update(self, net, lr, apply):
        """Gradient collective (+ optimizer step when `apply`).  The D update of an overlapped trainer runs on the side
        stream; whoever needs D's parameters next joins it (`_join_side`)."""
        self.sn_groups[net].backward()
        ar = self.store.arenas[net]
        for v in ar["vars"]:
            if not v.grad_set:
                v.grad_view.zero_()
        self._regularizer_grads(net, ar)
        side = apply and net == "discriminator" and self.comm_stream is not None
        if side:
            self.comm_stream.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(self.comm_stream):
                self._reduce_grads(net, ar, for_update=True)
                self._adam(net, lr)
            self._pending = True
            return
        self._reduce_grads(net, ar, for_update=apply)
        if apply:
            self._adam(net, lr)

    def _finish_grads(self, net, for_update=False):
        """(kept for callers that only want the summed gradients)"""
        self.sn_groups[net].backward()
        ar = self.store.arenas[net]
        for v in ar["vars"]:
            if not v.grad_set:
                v.grad_view.zero_()
        self._regularizer_grads(net, ar)
        self._reduce_grads(net, ar, for_update)
        return ar

    def _regularizer_grads(self, net, ar):
        self.reg_loss = None
        if net == "generator" and self.regs:
            # g_loss += sum of the regularisation losses (BigGAN.py:897-898).  Every rank holds the same weights and the
            # gradient arenas are summed across ranks afterwards, so each rank adds 1/world of the regulariser's gradient.
            self.reg_loss = torch.zeros(1, dtype=torch.float32, device=ar["grads"].device)
            for v, reg in self.regs:
                reg.evaluate(v.data, self.reg_loss, dW=v.grad_view, grad_scale=1.0 / self.world)

    def _reduce_grads(self, net, ar, for_update):
        if self.world > 1:
            import torch.distributed as dist
            if self.shard and for_update:                        # only this rank's slice of the summed gradient is needed
                with _lib.timed("nccl_reduce_scatter_grads"):
                    dist.reduce_scatter_tensor(self.opt[net]["gshard"], ar["grads"], group=self.pg)
            else:
                with _lib.timed("nccl_allreduce_grads"):
                    dist.all_reduce(ar["grads"], group=self.pg)      # NCCL over NVLink; sum == global-batch gradient

    def _adam(self, net, lr):
        """Adam(+EMA) over the network's flat arena; the step counter and lr_t live on the device (graph-replayable)."""
        cfg = self.cfg
        ar = self.store.arenas[net]
        o = self.opt[net]
        call("bg_adam_lr_step", o["state"], float(lr), float(cfg.beta1), float(cfg.beta2), stream())
        if self.shard:
            import torch.distributed as dist
            sl = self._shard_slice(net)
            p = o["pshard"]
            p.copy_(ar["params"][sl])
            call("bg_adam_ema_step", p, o["gshard"], o["m"][sl], o["v"][sl], o["ema"][sl] if o["ema"] is not None else None,
                 p.numel(), 0.0, o["state"][1:], float(cfg.beta1), float(cfg.beta2), 1e-8, float(cfg.moving_decay), 1.0, stream())
            with _lib.timed("nccl_allgather_params"):
                dist.all_gather_into_tensor(ar["params"], p, group=self.pg)
            return
        call("bg_adam_ema_step", ar["params"], ar["grads"], o["m"], o["v"], o["ema"], ar["params"].numel(),
             0.0, o["state"][1:], float(cfg.beta1), float(cfg.beta2), 1e-8, float(cfg.moving_decay), 1.0, stream())

    def full_opt(self, net, key):
        """Optimizer state `key` ('m', 'v', 'ema') over the whole arena (with a sharded optimizer every rank maintains its
        slice only: collective call)."""
        self._join_side()
        t = self.opt[net][key]
        if self.shard and t is not None:
            import torch.distributed as dist
            full = torch.empty_like(t)
            dist.all_gather_into_tensor(full, t[self._shard_slice(net)].clone(), group=self.pg)
            return full
        return t

    def full_ema(self):
        return self.full_opt("generator", "ema")

    def _as_input(self, img):
        t = Tensor(img)
        if ops.rgb_pad():
            return ops.pad_channels(t, ops.rgb_pad(), dtype=self.compute_dtype)     # cast + zero-pad 3 -> 16 channels
        return ops.cast(t, self.compute_dtype)

    # ------------------------------------------------------------------------------------------
    def d_update(self, real, labels, z, cls_z, apply=True):
        """d_train_op (BigGAN.py:1063): D(DiffAugment(real)), D(DiffAugment(G(z))), hinge + class loss."""
        cfg = self.cfg
        if not self._built:
            self.build()
        self._set_phase("discriminator")
        Bl = real.shape[0]
        with no_grad():                                   # no gradient flows into G in the D update
            fake = generator(cfg, z, cls_z)
            aug_real = DiffAugment_tf.DiffAugment(self._as_input(real), policy=cfg.da_policy)       # :806
            aug_fake = DiffAugment_tf.DiffAugment(fake, policy=cfg.da_policy)                       # :857
            x = Tensor(torch.cat([aug_real.data, aug_fake.data], dim=0))
            x.aux = aug_fake.aux
        if cfg.bn_in_d:
            # batch statistics inside D: the two instantiations (BigGAN.py:807, :857) normalise their own batch, so they
            # cannot share one pass.
            # Both instantiations move D's moving statistics with an assign_sub whose delta is formed from the value the
            # variable had when the run began (oracle/tf1_eager.py, session model): final = s_real + s_fake - start.
            moving = [v for k, v in self.store.vars.items() if "discriminator" in k and "/moving_" in k]
            start = [v.data.clone() for v in moving]
            out_r = discriminator(cfg, aug_real)
            after_real = [v.data.clone() for v in moving]
            for v, s0 in zip(moving, start):
                v.data.copy_(s0)
            out_f = discriminator(cfg, aug_fake)
            for v, s0, s1 in zip(moving, start, after_real):
                n = v.data.numel()
                call("bg_add", s1, v.data, s1, n, _lib.BG_F32, stream())                              # s_real + s_fake
                call("bg_axpby", s0, s1, None, -1.0, 1.0, v.data, n, _lib.BG_F32, stream())            # - start
            real_logits, fake_logits = out_r["real"], out_f["real"]
            out = {"cls": out_r.get("cls")}
        else:
            # D(real) and D(fake) share weights and have no batch statistics -> one pass over [real; fake]
            out = discriminator(cfg, x)
            real_logits, fake_logits = split_rows(out["real"], [Bl, Bl])
        d_loss = ops.discriminator_loss(cfg.gan_type, real=real_logits, fake=fake_logits, flood_level=cfg.d_flood)   # :879
        if cfg.acgan and cfg.bn_in_d:
            cls_loss = ops.cls_loss_fn('logistic', self.cls_weights)
            d_loss = d_loss + cfg.d_cls_loss_weight * cls_loss(labels, out["cls"])                 # :853
        elif cfg.acgan:
            real_cls, _ = split_rows(out["cls"], [Bl, Bl])
            cls_loss = ops.cls_loss_fn('logistic', self.cls_weights)
            d_loss = d_loss + cfg.d_cls_loss_weight * cls_loss(labels, real_cls)                   # :853
        backward([(d_loss, torch.ones(1, device=real.device))])
        self._update("discriminator", cfg.d_lr, apply)
        return d_loss.data, {"fake": fake, "real_logits": real_logits, "fake_logits": fake_logits}

    def g_update(self, z, cls_z, apply=True):
        """g_train_op (BigGAN.py:1082): fresh z, G -> DiffAugment -> D, -mean(fake) + class loss."""
        cfg = self.cfg
        if not self._built:
            self.build()
        self._set_phase("generator")
        fake = generator(cfg, z, cls_z)
        aug = DiffAugment_tf.DiffAugment(fake, policy=cfg.da_policy)
        self._sn_forward_discriminator()                   # joins the D update running underneath the generator forward
        out = discriminator(cfg, aug)
        g_loss = ops.generator_loss(cfg.gan_type, fake=out["real"], real=None, flood_level=cfg.g_flood)   # :896
        if cfg.acgan:
            cls_loss = ops.cls_loss_fn('logistic', self.cls_weights)
            g_loss = g_loss + cfg.g_cls_loss_weight * cls_loss(cls_z, out["cls"])                  # :894
        backward([(g_loss, torch.ones(1, device=z.device))])
        self._update("generator", cfg.g_lr, apply)
        loss = g_loss.data
        if self.reg_loss is not None:
            loss = torch.empty_like(g_loss.data)      # like the other terms, each rank reports its 1/world share: the sum over ranks is the loss
            call("bg_axpby", self.reg_loss, g_loss.data, None, 1.0 / self.world, 1.0, loss, 1, _lib.BG_F32, stream())
        return loss, {"fake": fake, "fake_logits": out["real"]}

    def train_step(self, real, labels, z_d, z_g, cls_z):
        """One iteration: D update, then G update (n_critic = 1, BigGAN.py:1080)."""
        d_loss, _ = self.d_update(real, labels, z_d, cls_z)
        g_loss, _ = self.g_update(z_g, cls_z)
        return d_loss, g_loss

    # ------------------------------------------------------------------------------------------
    # gradient accumulation over virtual batches, sampling with the EMA weights
    # ------------------------------------------------------------------------------------------
    def _accumulated_update(self, net, lr, run_one, batches):
        """utils.virtual_batch_steps (utils.py:242-250): zero the accumulators, one gradient evaluation per virtual batch
        (each its own run: spectral-norm `u` and BN statistics advance every time), then ONE optimizer step on acc / k."""
        k = len(batches)
        ar = self.store.arenas[net]
        n = ar["grads"].numel()
        acc = torch.zeros_like(ar["grads"])
        losses = []
        for b in batches:
            loss, _ = run_one(*b, apply=False)
            losses.append(loss)
            call("bg_add", acc, ar["grads"], acc, n, _lib.BG_F32, stream())
        call("bg_axpby", acc, None, None, 1.0 / k, 0.0, ar["grads"], n, _lib.BG_F32, stream())
        if self.shard:                                # the accumulated gradients are full sums (all-reduced per virtual batch)
            self.opt[net]["gshard"].copy_(ar["grads"][self._shard_slice(net)])
        self._adam(net, lr)
        return losses

    def d_update_virtual(self, batches):
        """batches: list of (real, labels, z, cls_z), len == cfg.virtual_batches."""
        return self._accumulated_update("discriminator", self.cfg.d_lr, self.d_update, batches)

    def g_update_virtual(self, batches):
        """batches: list of (z, cls_z)."""
        return self._accumulated_update("generator", self.cfg.g_lr, self.g_update, batches)

    def sample(self, z, cls_z, ema=True):
        """Generator in inference mode (is_training=False: population BN statistics, BigGAN.py:771-787) with the EMA
        shadow weights swapped in (MovingAverageOptimizer.swapping_saver semantics) when ema=True.  Returns the image
        tensor [B, S, S, 3(+pad)].  Spectral norm still advances `u` (ops.py:743 has no training switch).
        With a sharded optimizer and ema=True this is a COLLECTIVE call (the EMA shadow slices are all-gathered):
        every rank must make it."""
        if not self._built:
            self.build()
        self._join_side()
        ar, o = self.store.arenas["generator"], self.opt["generator"]
        saved = None
        if ema and o["ema"] is not None:
            saved = ar["params"].clone()
            ar["params"].copy_(self.full_ema())
        try:
            set_store(self.store)
            self.store.begin_phase()
            clear_tape()
            self.sn_groups["generator"].forward(self.store)
            with no_grad():
                img = generator(self.cfg, z, cls_z, is_training=False)
        finally:
            if saved is not None:
                ar["params"].copy_(saved)
        return img

    def grads(self, net):
        ar = self.store.arenas[net]
        self._join_side()
        torch.cuda.synchronize()
        return {v.name: v.grad_view.detach().cpu().numpy().copy() for v in ar["vars"]}
