"""Whole-step CUDA graph: the D update + G update of `Trainer.train_step` (several hundred kernel launches, NCCL
all-reduces included) captured once and replayed, so that the step is not launch-bound when the per-GPU batch is
small (data-parallel strong scaling) -- "CUDA streams and graphs instead of a tracing compiler".

Everything the step mutates lives on the device (weights, Adam state and step counters, SN `u`, BN running
statistics), the shapes are static and the host code has no data-dependent control flow, so a replay is exactly one
training iteration on the contents of the static input buffers."""
import torch

from .step import Trainer


class GraphedStep:
    def __init__(self, trainer: Trainer, real, labels, z_d, z_g, cls_z, warmup=2):
        if trainer.cfg.da_policy:
            raise ValueError("graph capture needs device-side DiffAugment draws; use the eager step with da_policy")
        self.tr = trainer
        self.static = [torch.empty_like(t).copy_(t) for t in (real, labels, z_d, z_g, cls_z)]
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):                       # warm-up off the default stream (allocator, NCCL, attributes)
            for _ in range(warmup):
                trainer.train_step(*self.static)
        torch.cuda.current_stream().wait_stream(side)
        torch.cuda.synchronize()
        from . import _lib
        n0 = _lib.launch_count
        self.graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(self.graph):
            self.d_loss, self.g_loss = trainer.train_step(*self.static)
        self.launches = _lib.launch_count - n0              # C-ABI calls captured in one step

    def __call__(self, real, labels, z_d, z_g, cls_z):
        for dst, src in zip(self.static, (real, labels, z_d, z_g, cls_z)):
            if dst.data_ptr() != src.data_ptr():
                dst.copy_(src, non_blocking=True)           # host (pinned) or device source
        self.graph.replay()
        return self.d_loss, self.g_loss
