"""Sample wire format: the byte framing SRL's sample streams carry (base/namedarray.py:100-218), and its decode
straight into the device buffer's pinned staging block.

A message is a list of byte strings: `[method code] + body + [pickle(metadata)]` (namedarray.py:163,213-216).
For the array methods the body is four frames per leaf -- key, dtype, shape, payload, all ASCII except the payload
(namedarray.py:115-128); a `None` leaf has empty dtype / shape / payload frames.

  method                        code    here
  pickle_dict                   0001    dumps / loads
  pickle                        0002    dumps / loads
  raw_bytes                     0003    dumps / loads / frames() / DeviceSlabBuffer.put_frames (zero intermediate copies)
  raw_compress, compress_pickle,
  pickle_compress, obs_compress,
  compress_except_policy_state  0004-8  need the third-party `blosc` codec (blosc.compress(typesize=4, cname='lz4'),
                                        namedarray.py:126), which is absent from this image and from /root/reference:
                                        its container format cannot be pinned to a single golden vector here, so these
                                        raise instead of guessing (the reference itself fails on `import blosc`).

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
_COMPRESSED = {b"0004": "raw_compress", b"0005": "compress_pickle", b"0006": "pickle_compress", b"0007": "obs_compress",
               b"0008": "compress_except_policy_state"}
_METHODS = {"pickle_dict": PICKLE_DICT, "pickle": PICKLE, "raw_bytes": RAW_BYTES}


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


def dumps(x: NamedArray, method: str = "pickle_dict") -> List[bytes]:
    """namedarray.py:100-163 for the methods this image can serve."""
    if "compress" in method:
        raise ModuleNotFoundError(f"method {method!r} needs the `blosc` codec, which is not installed")
    if method == "pickle_dict":
        body = [PICKLE_DICT, pickle.dumps((type(x).__name__, _to_dict(x)))]
    elif method == "pickle":
        body = [PICKLE, pickle.dumps(x)]
    elif method == "raw_bytes":
        body = [RAW_BYTES]
        for k, v in flatten(x):
            if v is None:
                body += [k.encode("ascii"), b"", b"", b""]
            else:
                v = np.asarray(v)
                body += [k.encode("ascii"), encode_dtype(v.dtype).encode("ascii"), str(tuple(v.shape)).encode("ascii"),
                         v.tobytes()]
    else:
        raise NotImplementedError(f"Unknown method {method}. Available are {sorted(_METHODS) + sorted(_COMPRESSED.values())}.")
    return body + [pickle.dumps(dict(**x.metadata))]


def frames(b: Sequence[bytes]) -> Tuple[List[Tuple[str, Optional[np.dtype], Optional[Tuple[int, ...]], Optional[memoryview]]], Dict]:
    """A raw_bytes message as [(dotted key, dtype, shape, payload view)] + metadata, without touching the payloads."""
    if bytes(b[0]) != RAW_BYTES:
        raise ValueError(f"frames() reads raw_bytes messages (code {RAW_BYTES!r}), got {bytes(b[0])!r}")
    xs = b[1:-1]
    if len(xs) % 4 != 0:
        raise ValueError(f"raw_bytes body has {len(xs)} frames, not a multiple of 4")
    out = []
    for i in range(len(xs) // 4):
        key = bytes(xs[4 * i]).decode("ascii")
        if len(xs[4 * i + 1]) == 0:  # namedarray.py:181-188: an empty dtype frame marks a None leaf
            out.append((key, None, None, None))
            continue
        dtype = np.dtype(bytes(xs[4 * i + 1]).decode("ascii"))
        shape = tuple(ast.literal_eval(bytes(xs[4 * i + 2]).decode("ascii")))
        payload = memoryview(xs[4 * i + 3]).cast("B")
        need = int(np.prod(shape, dtype=np.int64)) * dtype.itemsize
        if payload.nbytes != need:
            raise ValueError(f"leaf {key}: payload of {payload.nbytes} bytes, dtype {dtype} x shape {shape} needs {need}")
        out.append((key, dtype, shape, payload))
    return out, pickle.loads(bytes(b[-1]))


def loads(b: Sequence[bytes]) -> NamedArray:
    """namedarray.py:166-218: back to a NamedArray of (read-only, zero-copy) numpy arrays."""
    code = bytes(b[0])
    if code in _COMPRESSED:
        raise ModuleNotFoundError(f"method {_COMPRESSED[code]!r} needs the `blosc` codec, which is not installed")
    if code == PICKLE_DICT:
        _, values = pickle.loads(bytes(b[1]))
        x = from_dict(values)
        metadata = pickle.loads(bytes(b[-1]))
    elif code == PICKLE:
        x = pickle.loads(bytes(b[1]))
        metadata = pickle.loads(bytes(b[-1]))
    elif code == RAW_BYTES:
        entries, metadata = frames(b)
        x = from_flattened([(k, None if dt is None else np.frombuffer(p, dtype=dt).reshape(shape))
                            for k, dt, shape, p in entries])
    else:
        raise NotImplementedError(f"Unknown NamedArrayEncodingMethod value {code!r}.")
    x.metadata.clear()
    x.register_metadata(**metadata)
    return x
