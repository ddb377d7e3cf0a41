"""Sample wire format: the byte framing SRL's sample streams carry (base/namedarray.py:100-218), and its decode
straight into the device buffer's pinned staging block.

A message is a list of byte strings: `[method code] + body + [pickle(metadata)]` (namedarray.py:163,213-216).
For the array methods the body is four frames per leaf -- key, dtype, shape, payload, all ASCII except the payload
(namedarray.py:115-128); a `None` leaf has empty dtype / shape / payload frames.

  method                        code    body
  pickle_dict                   0001    pickle((class name, nested dict))
  pickle                        0002    pickle(object)
  raw_bytes                     0003    four frames per leaf
  raw_compress                  0004    four frames per leaf, every payload blosc-compressed
  compress_pickle               0005    pickle(list of the raw_compress frames)
  pickle_compress               0006    blosc(pickle(object))
  obs_compress                  0007    four frames per leaf, payloads of leaves with 'obs' in the key compressed
  compress_except_policy_state  0008    four frames per leaf, all payloads but those with 'policy_state' in the key compressed

The codec is the third-party `blosc` package (`blosc.compress(payload, typesize=4, cname='lz4')`, namedarray.py:126),
imported lazily exactly where the reference imports it (namedarray.py:101-103,168-171).  ENCODING always goes through it:
without it (this image, /root/reference) `dumps` of a compressed method raises ModuleNotFoundError, as the reference's does.
DECODING goes through the library's own decoder of the Blosc-1 / LZ4 frame (csrc/blosc_decode.cu: srl_blosc1_decompress) --
straight into the destination the caller names, which for `DeviceSlabBuffer.put_frames` is the pinned staging block: one pass
over the payload instead of blosc.decompress -> bytes -> copy -- chosen per process by `codec()`:
  * `blosc` importable: the native decoder is cross-checked ONCE against frames the real package writes (several sizes,
    compressible and incompressible payloads, the reference's typesize 4); it is used only if every one decodes identically,
    otherwise a warning is issued and the package decodes;
  * `blosc` not importable: the native decoder (its LZ4 layer is pinned to liblz4, its framing follows the published layout
    but is UNPINNED against blosc itself -- there is no blosc here to write vectors; csrc/blosc_decode.cu, tests/blosc1_writer.py);
  * `SRL_B200_WIRE_CODEC=blosc|native` overrides.
What is pinned against the reference is the FRAMING of the message -- which leaves are compressed, frame order, None leaves,
metadata -- from messages the unmodified reference wrote with a stand-in codec injected as `blosc` (tests/golden/wire.npz,
tests/test_wire.py).

The reference's `loads` allocates a fresh ndarray per leaf per message on the trainer's main thread
(distributed/system/sample_stream.py:176-198) and the buffer then copies every leaf again in `np.stack`.  `frames()`
only slices the message: each payload stays a memoryview of the received bytes, and `DeviceSlabBuffer.put_frames`
copies it once, into the pinned block that the H2D copy reads.
"""
from __future__ import annotations

import ast
import pickle
from typing import Any, Dict, List, Optional, Sequence, Tuple

import numpy as np

from srl_b200.namedarray import NamedArray, flatten, from_dict, from_flattened

PICKLE_DICT, PICKLE, RAW_BYTES = b"0001", b"0002", b"0003"
RAW_COMPRESS, COMPRESS_PICKLE, PICKLE_COMPRESS, OBS_COMPRESS, COMPRESS_EXCEPT_POLICY_STATE = (
    b"0004", b"0005", b"0006", b"0007", b"0008")
_METHODS = {"pickle_dict": PICKLE_DICT, "pickle": PICKLE, "raw_bytes": RAW_BYTES, "raw_compress": RAW_COMPRESS,
            "compress_pickle": COMPRESS_PICKLE, "pickle_compress": PICKLE_COMPRESS, "obs_compress": OBS_COMPRESS,
            "compress_except_policy_state": COMPRESS_EXCEPT_POLICY_STATE}
# which leaves of a framed message are compressed (namedarray.py:141-156,199-208)
_COMPRESS_IF = {RAW_BYTES: lambda k: False, RAW_COMPRESS: lambda k: True, COMPRESS_PICKLE: lambda k: True,
                OBS_COMPRESS: lambda k: "obs" in k, COMPRESS_EXCEPT_POLICY_STATE: lambda k: "policy_state" not in k}


def _blosc():
    import blosc  # third-party; raises ModuleNotFoundError where it is not installed, like namedarray.py:103,171
    return blosc


DECODE_THREADS = 1  # participants of one native decode (the frame's blocks are independent); DeviceSlabBuffer sets its own
_codec_choice: Optional[str] = None


def _native_decode_into(src, dst: np.ndarray, threads: int = 0) -> None:
    """One Blosc-1 frame -> `dst` (a contiguous uint8 array of exactly the frame's decoded size)."""
    from srl_b200 import _lib
    buf = np.frombuffer(src, dtype=np.uint8)
    _lib.call("srl_blosc1_decompress", buf.ctypes.data, buf.size, dst.ctypes.data if dst.size else None, dst.size,
              int(threads or DECODE_THREADS))


def _native_nbytes(src) -> int:
    import ctypes
    from srl_b200 import _lib
    buf = np.frombuffer(src, dtype=np.uint8)
    n = ctypes.c_size_t(0)
    _lib.call("srl_blosc1_info", buf.ctypes.data, buf.size, ctypes.byref(n), None, None, None, None)
    return int(n.value)


def _cross_check(blosc) -> bool:
    """The native decoder against frames the real package writes, once per process."""
    rng = np.random.default_rng(0)
    for n in (0, 1, 5, 127, 4096, 70_001, 1_300_003):
        for payload in (rng.integers(0, 256, n, dtype=np.uint8).tobytes(),
                        np.repeat(rng.integers(0, 256, n // 9 + 1, dtype=np.uint8), 9)[:n].tobytes()):
            frame = blosc.compress(payload, typesize=4, cname="lz4")
            out = np.empty(n, dtype=np.uint8)
            try:
                _native_decode_into(frame, out, 1)
            except Exception:  # noqa: BLE001
                return False
            if out.tobytes() != payload:
                return False
    return True


def codec() -> str:
    """'native' or 'blosc': who decodes compressed payloads in this process (see the module docstring)."""
    global _codec_choice
    if _codec_choice is not None:
        return _codec_choice
    import os
    import warnings
    want = os.environ.get("SRL_B200_WIRE_CODEC", "auto")
    if want not in ("auto", "native", "blosc"):
        raise ValueError(f"SRL_B200_WIRE_CODEC={want!r}: expected auto, native or blosc")
    choice = want
    if want == "auto":
        try:
            from srl_b200 import _lib
            _lib.load_library()
            native_ok = True
        except Exception:  # noqa: BLE001 -- no library (a host without the CUDA build): the package decodes, as in the reference
            native_ok = False
        try:
            blosc = _blosc()
        except ModuleNotFoundError:
            blosc = None
        if native_ok and blosc is not None:
            native_ok = _cross_check(blosc)
            if not native_ok:
                warnings.warn("srl_b200.wire: the native Blosc-1 decoder disagrees with the installed blosc package; "
                              "decoding with the package")
        choice = "native" if native_ok else "blosc"
    _codec_choice = choice
    return choice


def decompress(buf) -> bytes:
    """blosc.decompress(buf) by whoever `codec()` names."""
    if codec() == "blosc":
        return _blosc().decompress(bytes(buf))
    out = np.empty(_native_nbytes(buf), dtype=np.uint8)
    _native_decode_into(buf, out)
    return out.tobytes()


class CompressedLeaf:
    """A compressed payload that has not been decoded yet: `frames(b, lazy=True)` hands these out so that the buffer can
    decode each one straight into its pinned staging block (`decode_into`).  Behaves like a read-only array where one is
    asked for (`np.asarray(leaf)` decodes once and keeps the result)."""
    __slots__ = ("buf", "dtype", "shape", "nbytes", "_array")

    def __init__(self, buf, dtype: np.dtype, shape: Tuple[int, ...]):
        self.buf, self.dtype, self.shape = buf, np.dtype(dtype), tuple(shape)
        self.nbytes = int(np.prod(self.shape, dtype=np.int64)) * self.dtype.itemsize
        self._array = None
        have = _native_nbytes(buf)
        if have != self.nbytes:
            raise ValueError(f"compressed payload of {have} bytes, dtype {self.dtype} x shape {self.shape} needs {self.nbytes}")

    @property
    def size(self) -> int:
        return int(np.prod(self.shape, dtype=np.int64))

    @property
    def ndim(self) -> int:
        return len(self.shape)

    def decode_into(self, dst: np.ndarray, threads: int = 0) -> None:
        """`dst`: contiguous uint8 view of exactly `nbytes` bytes."""
        if dst.dtype != np.uint8 or dst.size != self.nbytes or not dst.flags.c_contiguous:
            raise ValueError(f"decode_into: need a contiguous uint8 destination of {self.nbytes} bytes")
        _native_decode_into(self.buf, dst, threads)

    def __array__(self, dtype=None, copy=None):
        if self._array is None:
            raw = np.empty(self.nbytes, dtype=np.uint8)
            self.decode_into(raw)
            self._array = raw.view(self.dtype).reshape(self.shape)
            self._array.flags.writeable = False
        return self._array if dtype is None else self._array.astype(dtype)


def encode_dtype(dtype) -> str:
    """base/numpy_utils.py:64-78 (bool travels as uint8)."""
    dtype = np.dtype(dtype)
    if dtype == np.uint8 or dtype == np.bool_:
        return "uint8"
    for name in ("float32", "float64", "int32", "int64"):
        if dtype == np.dtype(name):
            return name
    if str(dtype).startswith("<U"):
        return str(dtype)
    raise NotImplementedError(f"Data type to string not implemented: {dtype}.")


def _to_dict(x) -> Dict[str, Any]:
    return {k: (_to_dict(v) if isinstance(v, NamedArray) else v) for k, v in x.items()}


def _leaf_frames(x: NamedArray, compress_if) -> List[bytes]:
    """namedarray.py:112-128: key, dtype, shape, payload per leaf (empty frames for a None leaf)."""
    out: List[bytes] = []
    for k, v in flatten(x):
        if v is None:
            out += [k.encode("ascii"), b"", b"", b""]
            continue
        v = np.asarray(v)
        payload = v.tobytes()
        if compress_if(k):
            payload = _blosc().compress(payload, typesize=4, cname="lz4")
        out += [k.encode("ascii"), encode_dtype(v.dtype).encode("ascii"), str(tuple(v.shape)).encode("ascii"), payload]
    return out


def dumps(x: NamedArray, method: str = "pickle_dict") -> List[bytes]:
    """namedarray.py:100-163."""
    if method not in _METHODS:
        raise NotImplementedError(f"Unknown method {method}. Available are {sorted(_METHODS)}.")
    code = _METHODS[method]
    if "compress" in method:
        _blosc()  # fail before any work, as the reference does (namedarray.py:101-103)
    if code == PICKLE_DICT:
        body = [pickle.dumps((type(x).__name__, _to_dict(x)))]
    elif code == PICKLE:
        body = [pickle.dumps(x)]
    elif code == COMPRESS_PICKLE:
        body = [pickle.dumps(_leaf_frames(x, _COMPRESS_IF[code]))]
    elif code == PICKLE_COMPRESS:
        body = [_blosc().compress(pickle.dumps(x), typesize=4, cname="lz4")]
    else:
        body = _leaf_frames(x, _COMPRESS_IF[code])
    return [code] + body + [pickle.dumps(dict(**x.metadata))]


def frames(b: Sequence[bytes], lazy: bool = False):
    """A framed message (raw_bytes, raw_compress, compress_pickle, obs_compress, compress_except_policy_state) as
    [(dotted key, dtype, shape, payload view)] + metadata.  Uncompressed payloads stay views of the received bytes;
    compressed ones are decompressed once -- or, with `lazy` and the native decoder, handed out as `CompressedLeaf`s for the
    caller to decode where the bytes are needed."""
    code = bytes(b[0])
    if code not in _COMPRESS_IF:
        raise ValueError(f"frames() reads the framed methods {sorted(c.decode() for c in _COMPRESS_IF)}, got {code!r}")
    who = codec() if code != RAW_BYTES else None
    pkg = _blosc() if who == "blosc" else None  # namedarray.py:168-171: imported up front for codes 0004-0008
    xs = pickle.loads(bytes(b[1])) if code == COMPRESS_PICKLE else b[1:-1]
    if len(xs) % 4 != 0:
        raise ValueError(f"framed body has {len(xs)} frames, not a multiple of 4")
    compress_if = _COMPRESS_IF[code]
    out = []
    for i in range(len(xs) // 4):
        key = bytes(xs[4 * i]).decode("ascii")
        if len(xs[4 * i + 1]) == 0:  # namedarray.py:181-188: an empty dtype frame marks a None leaf
            out.append((key, None, None, None))
            continue
        dtype = np.dtype(bytes(xs[4 * i + 1]).decode("ascii"))
        shape = tuple(ast.literal_eval(bytes(xs[4 * i + 2]).decode("ascii")))
        buf = xs[4 * i + 3]
        if compress_if(key):
            if who == "native":
                leaf = CompressedLeaf(buf, dtype, shape)  # checks the decoded size against dtype x shape
                if lazy:
                    out.append((key, dtype, shape, leaf))
                    continue
                buf = np.asarray(leaf).reshape(-1).view(np.uint8)
            else:
                buf = pkg.decompress(bytes(buf))
        payload = memoryview(buf).cast("B")
        need = int(np.prod(shape, dtype=np.int64)) * dtype.itemsize
        if payload.nbytes != need:
            raise ValueError(f"leaf {key}: payload of {payload.nbytes} bytes, dtype {dtype} x shape {shape} needs {need}")
        out.append((key, dtype, shape, payload))
    return out, pickle.loads(bytes(b[-1]))


def loads(b: Sequence[bytes]) -> NamedArray:
    """namedarray.py:166-218: back to a NamedArray of (read-only, zero-copy where uncompressed) numpy arrays."""
    code = bytes(b[0])
    if code == PICKLE_DICT:
        _, values = pickle.loads(bytes(b[1]))
        x = from_dict(values)
    elif code == PICKLE:
        x = pickle.loads(bytes(b[1]))
    elif code == PICKLE_COMPRESS:
        x = pickle.loads(decompress(b[1]))
    elif code in _COMPRESS_IF:
        entries, _ = frames(b)
        x = from_flattened([(k, None if dt is None else np.frombuffer(p, dtype=dt).reshape(shape))
                            for k, dt, shape, p in entries])
    else:
        raise NotImplementedError(f"Unknown NamedArrayEncodingMethod value {code!r}.")
    metadata = pickle.loads(bytes(b[-1]))
    x.clear_metadata()  # namedarray.py:213-215
    x.register_metadata(**metadata)
    return x
