"""Peer-memory communicator for the small exchanges of the data-parallel step (csrc/comm.cu): the cross-replica
BatchNorm statistics are summed by ONE kernel that stores into every peer's buffer over NVLink and adds the slots in
rank order (bit-identical on all ranks), instead of an NCCL all-reduce per BN layer.  The CUDA IPC handles of the
exchange buffers travel through the caller's torch.distributed process group (plumbing)."""
import ctypes

import torch

from . import _lib
from ._lib import call
from .tensor import stream

MAX_DOUBLES = 4096


class PeerComm:
    def __init__(self, rank, world, process_group):
        import torch.distributed as dist
        self.rank, self.world = rank, world
        local = ctypes.c_void_p()
        handle = (ctypes.c_ubyte * 64)()
        call("bg_comm_local_alloc", ctypes.addressof(local), ctypes.addressof(handle))
        mine = torch.tensor(list(bytes(handle)), dtype=torch.uint8, device="cuda")
        every = torch.empty(world * 64, dtype=torch.uint8, device="cuda")
        dist.all_gather_into_tensor(every, mine, group=process_group)
        raw = bytes(every.cpu().numpy().tobytes())
        handles = (ctypes.c_ubyte * (64 * world)).from_buffer_copy(raw)
        comm = ctypes.c_void_p()
        call("bg_comm_open", rank, world, ctypes.addressof(handles), local, ctypes.addressof(comm))
        self.handle = comm
        dist.barrier(group=process_group)          # every rank has opened every buffer before the first exchange

    def all_reduce_sum(self, buf):
        """In-place sum over ranks of a float64 device vector (<= 4096 elements)."""
        assert buf.dtype == torch.float64 and buf.is_contiguous() and buf.numel() <= MAX_DOUBLES
        call("bg_comm_allreduce_f64", self.handle, buf, buf, buf.numel(), stream(), tag="p2p_allreduce_bn_stats")
        return buf

    def close(self):
        if self.handle is not None:
            _lib.lib.bg_comm_destroy(self.handle)
            self.handle = None
