"""Peer-memory exchange of the step's statistics table (srl_xchg_* in include/srl_b200.h).

`PeerExchange(group)` sets up the NVLink mailboxes of a one-node process group: every rank allocates its mailbox
through the C library, the 64-byte IPC handles travel once through `torch.distributed.all_gather_object` (host side),
and from then on `allreduce_sum(local, out)` is ONE kernel launch on the current stream -- no NCCL call, capturable in a
CUDA graph.  Reference: replaces the `dist.all_reduce` calls of legacy/algorithm/modules/utils.py:58-61,121-124.
"""
from __future__ import annotations

import ctypes

import torch
import torch.distributed as dist

from srl_b200 import _lib

HANDLE_BYTES = 64


class PeerExchange:

    def __init__(self, group, capacity_doubles: int, device: torch.device):
        self.world, self.rank = dist.get_world_size(group), dist.get_rank(group)
        if not 1 <= self.world <= 16:
            raise ValueError(f"PeerExchange serves 1..16 ranks of one node, got world size {self.world}")
        self.capacity = int(capacity_doubles)
        torch.cuda.set_device(device)
        h = ctypes.c_void_p()
        _lib.call("srl_xchg_create", self.world, self.rank, self.capacity, ctypes.byref(h))
        self._h = h
        mine = ctypes.create_string_buffer(HANDLE_BYTES)
        _lib.call("srl_xchg_local_handle", self._h, mine)
        gathered = [None] * self.world
        dist.all_gather_object(gathered, bytes(mine.raw), group=group)
        blob = ctypes.create_string_buffer(b"".join(gathered), HANDLE_BYTES * self.world)
        _lib.call("srl_xchg_connect", self._h, blob)
        dist.barrier(group=group)  # nobody launches before every mailbox is mapped everywhere

    def allreduce_sum(self, local: torch.Tensor, out: torch.Tensor) -> torch.Tensor:
        """out[i] = sum over ranks of local[i] (float64, contiguous CUDA tensors of <= capacity elements)."""
        for name, t in (("local", local), ("out", out)):
            if not (t.is_cuda and t.dtype == torch.float64 and t.is_contiguous()):
                raise ValueError(f"{name}: expected a contiguous CUDA float64 tensor")
        if local.numel() != out.numel() or local.numel() > self.capacity:
            raise ValueError(f"table of {local.numel()} doubles does not fit the exchange capacity {self.capacity}")
        _lib.call("srl_xchg_allreduce_sum", self._h, local.data_ptr(), out.data_ptr(), local.numel(),
                  torch.cuda.current_stream().cuda_stream)
        return out

    def check(self) -> None:
        """Synchronises and raises if a wait ever timed out (a peer did not launch its side of an exchange)."""
        st = ctypes.c_int(0)
        _lib.call("srl_xchg_status", self._h, ctypes.byref(st))
        if st.value != 0:
            raise RuntimeError("peer exchange timed out: a rank did not take part in a statistics exchange")

    def close(self) -> None:
        if getattr(self, "_h", None) is not None:
            _lib.call("srl_xchg_destroy", self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
