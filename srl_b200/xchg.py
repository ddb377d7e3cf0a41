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


class PeerExchangeUnavailable(RuntimeError):
    """Raised on EVERY rank of the group when any rank could not create or map the mailboxes."""


class PeerExchange:

    def __init__(self, group, capacity_doubles: int, device: torch.device, timeout_s: float = None):
        """Collective over `group`: every rank makes the same calls in the same order whatever fails locally, and the
        outcome is agreed on (a rank whose cudaIpcOpenMemHandle fails must not leave the others in the mailbox protocol
        while it falls back to NCCL): either all ranks return a connected exchange or all raise PeerExchangeUnavailable."""
        self.world, self.rank = dist.get_world_size(group), dist.get_rank(group)
        if not 1 <= self.world <= 16:
            raise ValueError(f"PeerExchange serves 1..16 ranks of one node, got world size {self.world}")
        self.capacity = int(capacity_doubles)
        self._h = None
        self._status_pin = torch.zeros(1, dtype=torch.int32).pin_memory()
        err, mine = None, b""
        try:
            torch.cuda.set_device(device)
            h = ctypes.c_void_p()
            _lib.call("srl_xchg_create", self.world, self.rank, self.capacity, ctypes.byref(h))
            self._h = h
            buf = ctypes.create_string_buffer(HANDLE_BYTES)
            _lib.call("srl_xchg_local_handle", self._h, buf)
            mine = bytes(buf.raw)
        except Exception as e:  # noqa: BLE001 -- reported to the group below
            err = f"rank {self.rank}: {e}"
        gathered = [None] * self.world
        dist.all_gather_object(gathered, (err, mine), group=group)
        errs = [g[0] for g in gathered if g[0] is not None]
        if not errs:
            try:
                blob = ctypes.create_string_buffer(b"".join(g[1] for g in gathered), HANDLE_BYTES * self.world)
                _lib.call("srl_xchg_connect", self._h, blob)
                if timeout_s is not None:
                    self.set_timeout(timeout_s)
            except Exception as e:  # noqa: BLE001
                err = f"rank {self.rank}: {e}"
            second = [None] * self.world
            dist.all_gather_object(second, err, group=group)  # also the barrier: nobody launches before every mailbox is mapped
            errs = [g for g in second if g is not None]
        if errs:
            self.close()
            raise PeerExchangeUnavailable("; ".join(errs))

    def set_timeout(self, seconds: float) -> None:
        """How long a rank waits for its peers (default ~10 minutes); applies to launches / graph captures made afterwards."""
        _lib.call("srl_xchg_set_timeout", self._h, float(seconds))

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

    def check_async(self) -> None:
        """Queues a copy of the status word into pinned memory on the current stream; `raise_if_failed()` reads it after
        the caller's own synchronisation (the trainer's one sync per step)."""
        _lib.call("srl_xchg_status_async", self._h, self._status_pin.data_ptr(), torch.cuda.current_stream().cuda_stream)

    def raise_if_failed(self) -> None:
        if int(self._status_pin[0]) != 0:
            raise RuntimeError("peer exchange timed out: a rank did not take part in a statistics exchange; the "
                               "statistics of this step are NaN by construction and must not be used")

    def close(self) -> None:
        if getattr(self, "_h", None) is not None:
            _lib.call("srl_xchg_destroy", self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
