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
imported lazily exactly where the reference imports it (namedarray.py:101-103,168-171): with blosc installed (any SRL
deployment that uses these methods) all eight methods work; without it (this image, /root/reference) the compressed
ones raise ModuleNotFoundError, as the reference's do.  What is pinned here is the FRAMING -- which leaves are compressed,
frame order, None leaves, metadata -- against messages the unmodified reference wrote with a stand-in codec injected as
`blosc` (tests/golden/wire.npz, tests/test_wire.py); the codec's own byte format never passes through this file's logic.

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


def frames(b: Sequence[bytes]) -> Tuple[List[Tuple[str, Optional[np.dtype], Optional[Tuple[int, ...]], Optional[memoryview]]], Dict]:
    """A framed message (raw_bytes, raw_compress, compress_pickle, obs_compress, compress_except_policy_state) as
    [(dotted key, dtype, shape, payload view)] + metadata.  Uncompressed payloads stay views of the received bytes;
    compressed ones are decompressed once (needs blosc)."""
    code = bytes(b[0])
    if code not in _COMPRESS_IF:
        raise ValueError(f"frames() reads the framed methods {sorted(c.decode() for c in _COMPRESS_IF)}, got {code!r}")
    codec = _blosc() if code != RAW_BYTES else None  # namedarray.py:168-171: imported up front for codes 0004-0008
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
            buf = codec.decompress(bytes(buf))
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
        x = pickle.loads(_blosc().decompress(bytes(b[1])))
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
