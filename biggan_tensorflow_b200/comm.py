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
    """Construction is collective and never skips a collective on a local failure: every rank reports its status with
    the handles, and `ok` (identical on all ranks after `agree()`) says whether the communicator may be used."""

    def __init__(self, rank, world, process_group):
        import torch.distributed as dist
        self.rank, self.world, self.handle, self.error = rank, world, None, None
        local = ctypes.c_void_p()
        handle = (ctypes.c_ubyte * 64)()
        good = 1
        try:
            call("bg_comm_local_alloc", ctypes.addressof(local), ctypes.addressof(handle))
        except Exception as e:               # noqa: BLE001  (reported to the peers below)
            good, self.error = 0, e
        mine = torch.tensor(list(bytes(handle)) + [good], dtype=torch.uint8, device="cuda")
        every = torch.empty(world * 65, dtype=torch.uint8, device="cuda")
        dist.all_gather_into_tensor(every, mine, group=process_group)
        rows = every.cpu().numpy().reshape(world, 65)
        if good and int(rows[:, 64].min()) == 1:
            handles = (ctypes.c_ubyte * (64 * world)).from_buffer_copy(rows[:, :64].tobytes())
            comm = ctypes.c_void_p()
            try:
                call("bg_comm_open", rank, world, ctypes.addressof(handles), local, ctypes.addressof(comm))
                self.handle = comm
            except Exception as e:           # noqa: BLE001
                self.error = e
        self.ok = self.handle is not None
        self._pg = process_group

    def agree(self):
        """Collective: True iff EVERY rank opened every peer buffer (doubles as the barrier before the first exchange)."""
        import torch.distributed as dist
        flag = torch.tensor([1.0 if self.ok else 0.0], device="cuda")
        dist.all_reduce(flag, op=dist.ReduceOp.MIN, group=self._pg)
        self.ok = float(flag.cpu()) > 0.5
        return self.ok

    def all_reduce_sum(self, buf):
        """In-place sum over ranks of a float64 device vector (<= 4096 elements)."""
        assert buf.dtype == torch.float64 and buf.is_contiguous() and buf.numel() <= MAX_DOUBLES
        call("bg_comm_allreduce_f64", self.handle, buf, buf, buf.numel(), stream(), tag="p2p_allreduce_bn_stats")
        return buf

    def close(self):
        if self.handle is not None:
            _lib.lib.bg_comm_destroy(self.handle)
            self.handle = None
