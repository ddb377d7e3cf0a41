"""Device-resident sample buffer: drop-in for SRL's `PriorityQueueBuffer` (base/buffer.py:87-165) whose batches
live in HBM.

Reference behaviour kept (base/buffer.py:109-165, pinned by base/tests/buffer_test.py:44-73):
  * `put(x)` stamps `trainer_worker_recv_timestamp`, collects `batch_size` samples and turns them into one batch with
    `recursive_aggregate(samples, np.stack(axis=1))` (leaves `[L, ...] -> [L, B, ...]`, `None` leaves zero-filled when
    only some samples miss them, base/namedarray.py:588-633); returns True when a batch was formed;
  * entries are ordered by (reuses_left, receive_time); `get()` serves the largest (most reuses left, newest = LIFO),
    decrements `reuses_left`, and re-inserts the entry while it has reuses left and the buffer is not full;
  * more than `max_size` entries drop the smallest; `get()` on an empty buffer raises AssertionError;
  * `batch_size == 0`: no batching, objects pass through untouched.

What changes is where the bytes go.  The reference stacks on the host with single-threaded `np.stack` (~1 GB/s,
SURVEY.md App. C) and the trainer then copies every leaf to the device and inflates it to float32
(api/trainer.py:215-217).  Here every incoming sample is packed into ONE pinned staging block and sent with ONE
asynchronous H2D copy into a device staging slot (whole samples one after another, `[slot][leaf][L, row]`); when a
batch is complete one `srl_batch_gather` call performs the stack -- `dst[t, j, :] = stage[slot_j][t, :]` for every leaf
at once -- into contiguous `[L, B, ...]` device tensors in the samples' own dtypes (uint8 frames stay uint8).  The batch
the trainer receives is a NamedArray of CUDA tensors; `MultiAgentPPOB200.step` consumes it without any host copy, and
a re-served entry carries its cached `adv` / `ret` on the device (mappo.py:224-225,254-257).
"""
from __future__ import annotations

import bisect
import dataclasses
import time
from typing import Any, Dict, List, Optional, Tuple

import numpy as np
import torch

from srl_b200 import ops, wire
from srl_b200._lib import LeafDesc
from srl_b200.namedarray import flatten, from_flattened

_MT_COPY_MIN = 1 << 20  # leaves at least this large go through the multi-threaded host copy (copy_threads > 1)
_ALIGN = 256  # leaf offsets inside a staged sample (keeps every leaf 16-byte aligned for 128-bit copies)


@dataclasses.dataclass(order=True)
class ReplayEntry:
    """base/buffer.py:24-34."""
    reuses_left: int
    receive_time: float
    sample: Any = dataclasses.field(compare=False)
    reuses: int = dataclasses.field(default=0, compare=False)
    sampling_indices: Optional[np.ndarray] = None

    def __len__(self):
        return len(self.sample)


def _leaf(v):
    """A leaf as the staging code looks at it (dtype, shape, nbytes, size): an ndarray, a still-compressed wire payload (kept
    as it is -- _stage decodes it straight into the pinned block) or anything array-like."""
    return v if isinstance(v, (np.ndarray, wire.CompressedLeaf)) else np.asarray(v)


class _Layout:
    """Byte layout of one staged sample: leaf name -> (offset, dtype, shape [L, ...]); fixed by the first sample."""

    def __init__(self, leaves: List[Tuple[str, Optional[np.ndarray]]]):
        self.names = [k for k, _ in leaves]
        self.spec: Dict[str, Tuple[int, np.dtype, Tuple[int, ...]]] = {}
        off = 0
        self.L = None
        for k, v in leaves:
            if v is None:
                continue
            v = _leaf(v)
            if self.L is None:
                self.L = v.shape[0]
            if v.shape[0] != self.L:
                raise ValueError(f"leaf {k}: leading dim {v.shape[0]} differs from the other leaves ({self.L})")
            self.spec[k] = (off, v.dtype, tuple(v.shape))
            off += (v.nbytes + _ALIGN - 1) // _ALIGN * _ALIGN
        self.bytes = max(off, _ALIGN)

    def matches(self, k: str, v: np.ndarray) -> bool:
        s = self.spec.get(k)
        return s is not None and s[1] == v.dtype and s[2] == tuple(v.shape)

    def extended_with(self, leaves) -> "_Layout":
        """Layout that also holds the leaves this one has not seen yet (a leaf that was None in earlier samples)."""
        example: Dict[str, Optional[np.ndarray]] = {k: None for k in self.names}
        for k, (_, dtype, shape) in self.spec.items():
            example[k] = np.empty(shape, dtype=dtype)
        for k, v in leaves:
            if example.get(k) is None:
                example[k] = None if v is None else _leaf(v)
        return _Layout(sorted(example.items(), key=lambda kv: kv[0]))


class DeviceSlabBuffer:
    """`make_buffer("priority_queue", ...)` with device-resident batches (see the module docstring)."""

    def __init__(self, max_size: int = 16, reuses: int = 1, batch_size: int = 1, device=None, staging_batches: int = 2,
                 copy_threads: int = 1):
        if not torch.cuda.is_available():
            raise RuntimeError("DeviceSlabBuffer needs a CUDA device (srl_b200 has no CPU path; use SRL's own "
                               "base.buffer.PriorityQueueBuffer on CPU)")
        ops._lib.load_library()
        self.__buffer: List[ReplayEntry] = []
        self.__max_size = max_size
        self.reuses = reuses
        self.batch_size = batch_size
        self.device = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
        self._layout: Optional[_Layout] = None
        self._n_slots = max(1, batch_size) * max(1, staging_batches)
        self._free: List[int] = []
        self._pending: List[Tuple[int, set, Any]] = []  # (slot, names present, (leaves, metadata)) of the next batch
        self._copy_stream = torch.cuda.Stream(device=self.device)
        self._slot_ready: Dict[int, torch.cuda.Event] = {}
        self._slot_free_after: Dict[int, torch.cuda.Event] = {}
        self.bytes_staged = 0
        # host threads that copy a large leaf into the pinned block (srl_host_copy); 1 = numpy's own single copy
        self.copy_threads = max(1, int(copy_threads))

    # ---- reference interface -------------------------------------------------------------------------------------
    @property
    def overflow(self):
        return len(self.__buffer) > self.__max_size

    def full(self):
        return len(self.__buffer) == self.__max_size

    def empty(self):
        return len(self.__buffer) == 0

    def qsize(self):
        return len(self.__buffer)

    def put(self, x) -> bool:
        if not self.batch_size:  # base/buffer.py:128-129
            self.__put(ReplayEntry(reuses_left=self.reuses, sample=x, receive_time=time.time()))
            return False
        x.trainer_worker_recv_timestamp = np.full(shape=x.on_reset.shape, fill_value=int(time.time()),
                                                  dtype=np.int64)  # base/buffer.py:114-116
        self._stage(flatten(x), getattr(x, "metadata", None))
        return self._batch_if_complete()

    def put_frames(self, b) -> bool:
        """`put(namedarray.loads(b))` for a framed message (raw_bytes, raw_compress, compress_pickle, obs_compress,
        compress_except_policy_state; base/namedarray.py:115-128,166-218) without building the sample: every leaf's payload
        is written once into the pinned staging block -- copied from the received bytes, or, for a compressed leaf, decoded
        there by the library's own Blosc-1 / LZ4 decoder (srl_blosc1_decompress; `wire.codec()` says when the `blosc`
        package decodes instead)."""
        if not self.batch_size:
            return self.put(wire.loads(b))
        entries, metadata = wire.frames(b, lazy=True)  # compressed payloads stay compressed until _stage decodes them
        leaves = [(k, None if dt is None else (p if isinstance(p, wire.CompressedLeaf) else
                                               np.frombuffer(p, dtype=dt).reshape(shape))) for k, dt, shape, p in entries]
        shape = next(v.shape for k, v in leaves if k == "on_reset")
        stamp = np.full(shape=shape, fill_value=int(time.time()), dtype=np.int64)  # base/buffer.py:114-116
        leaves = [(k, stamp if k == "trainer_worker_recv_timestamp" else v) for k, v in leaves]
        if not any(k == "trainer_worker_recv_timestamp" for k, _ in leaves):
            leaves = sorted(leaves + [("trainer_worker_recv_timestamp", stamp)], key=lambda kv: kv[0])
        self._stage(leaves, metadata)
        return self._batch_if_complete()

    def _batch_if_complete(self) -> bool:
        if len(self._pending) >= self.batch_size:
            data = self._assemble(self._pending[:self.batch_size])
            self._pending = self._pending[self.batch_size:]
            self.__put(ReplayEntry(reuses_left=self.reuses, sample=data, receive_time=time.time()))
            return True
        return False

    def get(self) -> ReplayEntry:
        assert not self.empty(), "attempting to get from empty buffer."
        r = self.__buffer.pop(-1)
        r.reuses_left -= 1
        r.reuses += 1
        if not self.full() and r.reuses_left > 0:
            self.__put(r)
        return r

    def __put(self, r: ReplayEntry) -> None:
        bisect.insort(self.__buffer, r)
        while self.overflow:
            self.__buffer.pop(0)  # base/buffer.py:164-165

    # ---- staging ----------------------------------------------------------------------------------------------------
    def _init_layout(self, leaves) -> None:
        self._layout = _Layout(leaves)
        lay = self._layout
        self._stage_dev = torch.empty((self._n_slots, lay.bytes), dtype=torch.uint8, device=self.device)
        self._stage_pin = torch.empty((self._n_slots, lay.bytes), dtype=torch.uint8).pin_memory()
        self._free = list(range(self._n_slots))

    def _relayout(self, leaves) -> None:
        torch.cuda.synchronize(self.device)
        pending = self._pending
        self._layout = self._layout.extended_with(leaves)
        lay = self._layout
        self._stage_dev = torch.empty((self._n_slots, lay.bytes), dtype=torch.uint8, device=self.device)
        self._stage_pin = torch.empty((self._n_slots, lay.bytes), dtype=torch.uint8).pin_memory()
        self._free = list(range(self._n_slots))
        self._slot_ready.clear()
        self._slot_free_after.clear()
        self._pending = []
        for _, _, (old_leaves, old_meta), _ in pending:
            self._stage(old_leaves, old_meta)

    @staticmethod
    def _host_only(v) -> bool:
        """Leaves that cannot (or need not) live on the device: non-numeric dtypes -- real SRL samples carry `policy_name`
        as a '<U..' string array (policy_worker.py:186; the trainer worker clears it after get(), trainer_worker.py:169) --
        and zero-size leaves.  They stay on the host and are stacked with np.stack, as the reference does for every leaf."""
        v = _leaf(v)
        return v.dtype.kind not in "fiub" or v.size == 0

    def _stage(self, leaves, metadata) -> None:
        """One sample (its flattened leaves) -> one pinned block -> one async H2D into a device staging slot."""
        all_leaves = leaves
        host_only = {k: np.asarray(v) for k, v in leaves if v is not None and self._host_only(v)}
        if host_only:
            leaves = [(k, None if k in host_only else v) for k, v in leaves]
        if self._layout is None:
            self._init_layout(leaves)
        elif any(v is not None and k not in self._layout.spec for k, v in leaves):
            self._relayout(leaves)  # a leaf that was None so far: rare, re-stages the pending samples
        lay = self._layout
        if not self._free:
            raise RuntimeError("DeviceSlabBuffer: no free staging slot (more than staging_batches * batch_size samples "
                               "pending); raise staging_batches")
        slot = self._free.pop(0)
        ev = self._slot_free_after.pop(slot, None)
        if ev is not None:
            ev.synchronize()  # the gather that last read this slot has finished
        pin = self._stage_pin[slot].numpy()
        present = set()
        for k, v in leaves:
            if v is None:
                continue
            v = _leaf(v)
            if not lay.matches(k, v):
                raise ValueError(f"leaf {k}: dtype/shape {v.dtype}{v.shape} differs from the first sample's "
                                 f"{lay.spec.get(k, ('-', None, None))[1:]} (samples of one buffer share their layout)")
            off = lay.spec[k][0]
            if isinstance(v, wire.CompressedLeaf):  # Blosc-1 / LZ4 frame -> pinned block, one pass (csrc/blosc_decode.cu)
                v.decode_into(pin[off:off + v.nbytes], self.copy_threads)
            elif self.copy_threads > 1 and v.nbytes >= _MT_COPY_MIN:
                src = np.ascontiguousarray(v)
                ops._lib.call("srl_host_copy", pin.ctypes.data + off, src.ctypes.data, src.nbytes, self.copy_threads)
            else:
                pin[off:off + v.nbytes] = np.ascontiguousarray(v).view(np.uint8).reshape(-1)
            present.add(k)
        with torch.cuda.stream(self._copy_stream):
            self._stage_dev[slot].copy_(self._stage_pin[slot], non_blocking=True)
            done = torch.cuda.Event()
            done.record(self._copy_stream)
        self._slot_ready[slot] = done
        self.bytes_staged += lay.bytes
        self._pending.append((slot, present, (all_leaves, metadata), host_only))

    def _assemble(self, items) -> Any:
        """np.stack(axis=1) of the staged samples, every leaf in one srl_batch_gather call."""
        lay = self._layout
        B = len(items)
        slots = [it[0] for it in items]
        main = torch.cuda.current_stream()
        for s in slots:
            main.wait_event(self._slot_ready.pop(s))
        idx = torch.tensor(slots, dtype=torch.int32, device=self.device)
        out: Dict[str, Optional[torch.Tensor]] = {}
        descs = []
        names = [k for k in lay.names if k in lay.spec or not any(n.startswith(k + ".") for n in lay.spec)]
        for k in names:
            host = [it[3].get(k) for it in items]
            if any(h is not None for h in host):  # host-only leaf: np.stack(axis=1), zero-filled where a sample had None
                like = next(h for h in host if h is not None)
                out[k] = np.stack([np.zeros_like(like) if h is None else h for h in host], axis=1)
                continue
            if k not in lay.spec or not any(k in it[1] for it in items):
                out[k] = None  # None in every sample stays None (base/namedarray.py:610)
                continue
            off, dtype, shape = lay.spec[k]
            row_bytes = int(np.prod(shape[1:], dtype=np.int64)) * dtype.itemsize
            dst = torch.empty((lay.L, B) + shape[1:], dtype=_torch_dtype(dtype), device=self.device)
            missing = [j for j, it in enumerate(items) if k not in it[1]]
            descs.append(LeafDesc(self._stage_dev.data_ptr() + off, dst.data_ptr(), row_bytes, self._n_slots, row_bytes,
                                  lay.bytes))
            out[k] = dst
            if missing:  # zero-filled where a sample had None (base/namedarray.py:588-595)
                out[k] = (dst, missing)
        ops.stack_samples(descs, idx, lay.L, B)
        for k, v in list(out.items()):
            if isinstance(v, tuple):
                v[0][:, v[1]] = 0
                out[k] = v[0]
        done = torch.cuda.Event()
        done.record(main)
        for s in slots:
            self._slot_free_after[s] = done
            self._free.append(s)
        batch = from_flattened([(k, out[k]) for k in names])
        meta = items[0][2][1]
        if meta is not None and hasattr(batch, "register_metadata"):
            batch.register_metadata(**meta)
        return batch


def _torch_dtype(dt: np.dtype) -> torch.dtype:
    return torch.from_numpy(np.empty(0, dtype=dt)).dtype


def make_buffer(name: str, **buffer_args):
    """base/buffer.py:533-541 for the one buffer kind on the PPO trainer path."""
    if name in ("priority_queue", "device_priority_queue"):
        return DeviceSlabBuffer(**buffer_args)
    raise NotImplementedError(f"srl_b200.buffer only provides the trainer-side priority queue, not {name!r}")
