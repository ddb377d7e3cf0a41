"""CPU tests of the library's own decoder of compressed sample-wire payloads (csrc/blosc_decode.cu: srl_lz4_block_decompress,
srl_blosc1_decompress) and of its use by srl_b200/wire.py.

What pins what: the LZ4 layer decodes blocks written by liblz4 itself (pyarrow's `lz4_raw` codec).  The Blosc-1 framing is
checked against tests/blosc1_writer.py, a writer that follows the published layout -- `blosc` is not installed here, so the
framing is UNPINNED against the real package (wire.codec() cross-checks it at run time where the package exists)."""
import ctypes
import sys
import types

import numpy as np
import pytest

from srl_b200 import _lib, wire
from srl_b200.api import AnalyzedResult, SampleBatch
from srl_b200.namedarray import NamedArray, flatten
from tests import blosc1_writer as W


def _payloads(n, rng):
    yield "random", rng.integers(0, 256, n, dtype=np.uint8).tobytes()
    yield "runs", np.repeat(rng.integers(0, 256, n // 7 + 1, dtype=np.uint8), 7)[:n].tobytes()
    yield "sparse", (rng.integers(0, 4, n) * rng.integers(0, 2, n)).astype(np.uint8).tobytes()
    yield "period", np.tile(rng.integers(0, 256, 1 + n % 23, dtype=np.uint8), n // (1 + n % 23) + 1)[:n].tobytes()
    yield "floats", np.cumsum(rng.standard_normal(n // 4 + 1)).astype(np.float32).tobytes()[:n]


def _lz4(block: bytes, n: int, cap=None) -> bytes:
    cap = n if cap is None else cap
    out = np.empty(max(cap, 1), dtype=np.uint8)
    src = np.frombuffer(block, dtype=np.uint8)
    w = ctypes.c_size_t(0)
    _lib.call("srl_lz4_block_decompress", src.ctypes.data, len(block), out.ctypes.data, cap, ctypes.byref(w))
    return out[:w.value].tobytes()


def _blosc1(frame: bytes, n: int, threads: int = 1) -> bytes:
    out = np.empty(n, dtype=np.uint8)
    src = np.frombuffer(frame, dtype=np.uint8)
    _lib.call("srl_blosc1_decompress", src.ctypes.data, len(frame), out.ctypes.data if n else None, n, threads)
    return out.tobytes()


@pytest.mark.parametrize("n", [0, 1, 4, 15, 16, 17, 64, 255, 4096, 65537, 1_000_003])
def test_lz4_blocks_written_by_liblz4_decode_bit_exact(n):
    rng = np.random.default_rng(n)
    for kind, payload in _payloads(n, rng):
        block = W.lz4_block(payload)
        assert _lz4(block, n) == payload, kind
        assert _lz4(block, n, cap=n + 100) == payload, kind  # a roomier destination: the same bytes, the same count


def test_lz4_malformed_blocks_are_errors_not_overruns():
    payload = np.repeat(np.arange(200, dtype=np.uint8), 9).tobytes()
    block = W.lz4_block(payload)
    with pytest.raises(_lib.SrlCudaError):
        _lz4(block, len(payload), cap=len(payload) - 1)  # does not fit
    with pytest.raises(_lib.SrlCudaError):
        _lz4(block[:-3], len(payload))  # cut inside the last literals
    with pytest.raises(_lib.SrlCudaError):
        _lz4(b"\x0f\x01\x00", 64)  # a match before any output: offset beyond the start
    with pytest.raises(_lib.SrlCudaError):
        _lz4(b"\x10A\x00\x00", 64)  # offset 0
    rng = np.random.default_rng(1)
    for _ in range(300):  # random bytes: an error or some output, never a crash
        junk = rng.integers(0, 256, rng.integers(1, 80), dtype=np.uint8).tobytes()
        try:
            _lz4(junk, 256)
        except _lib.SrlCudaError:
            pass


@pytest.mark.parametrize("typesize", [1, 2, 3, 4, 8])
@pytest.mark.parametrize("shuffle,split", [(True, True), (True, False), (False, True), (False, False)])
def test_blosc1_frames_decode_bit_exact(typesize, shuffle, split):
    rng = np.random.default_rng(typesize * 4 + shuffle * 2 + split)
    for n in (0, 1, 3, 4, 100, 511, 512, 513, 4096, 70_000, 300_001):
        for blocksize in (0, 512, 4096, 1 << 18):
            for kind, payload in _payloads(n, rng):
                frame = W.compress(payload, typesize, blocksize, shuffle, split)
                assert _blosc1(frame, n) == payload, (n, blocksize, kind)
    big = next(p for k, p in _payloads(3_000_001, rng) if k == "runs")  # several blocks, several threads
    assert _blosc1(W.compress(big, typesize, 1 << 18, shuffle, split), len(big), threads=4) == big


def test_blosc1_stored_frames_header_and_rejections():
    rng = np.random.default_rng(3)
    payload = rng.integers(0, 256, 5000, dtype=np.uint8).tobytes()
    assert _blosc1(W.compress(payload, memcpyed=True), 5000) == payload
    frame = W.compress(payload, 4, 1024)
    src = np.frombuffer(frame, dtype=np.uint8)
    nb, cb, bs, ts, fl = ctypes.c_size_t(), ctypes.c_size_t(), ctypes.c_size_t(), ctypes.c_int(), ctypes.c_int()
    _lib.call("srl_blosc1_info", src.ctypes.data, len(frame), ctypes.byref(nb), ctypes.byref(cb), ctypes.byref(bs),
              ctypes.byref(ts), ctypes.byref(fl))
    assert (nb.value, cb.value, bs.value, ts.value) == (5000, len(frame), 1024, 4) and fl.value & W.SHUFFLE
    bad = bytearray(frame)
    bad[2] = (bad[2] & 0x1F) | (4 << 5)  # codec family zstd
    with pytest.raises(_lib.SrlCudaError, match="codec family"):
        _blosc1(bytes(bad), 5000)
    bad = bytearray(frame)
    bad[2] |= 0x04  # bit-shuffle
    with pytest.raises(_lib.SrlCudaError, match="bit-shuffled"):
        _blosc1(bytes(bad), 5000)
    with pytest.raises(_lib.SrlCudaError):
        _blosc1(frame, 4999)  # destination of the wrong size
    with pytest.raises(_lib.SrlCudaError):
        _blosc1(frame[:40], 5000)  # header promises more bytes than there are
    for cut in range(17, len(frame), 97):  # a frame whose tail is garbage: an error, never a crash
        broken = frame[:cut] + bytes(len(frame) - cut)
        try:
            _blosc1(broken, 5000)
        except _lib.SrlCudaError:
            pass


def _sample(rng, L=9):
    frame = np.repeat(rng.integers(0, 256, (L, 4, 21, 84), dtype=np.uint8), 4, axis=2)  # image-like: compressible
    return SampleBatch(obs=NamedArray(frame=frame, vec=rng.standard_normal((L, 7)).astype(np.float32)),
                       on_reset=rng.integers(0, 2, (L, 1)).astype(np.uint8), done=np.zeros((L, 1), dtype=np.uint8), truncated=None,
                       action=NamedArray(x=rng.integers(0, 18, (L, 1)).astype(np.int32)),
                       reward=rng.standard_normal((L, 1)).astype(np.float32),
                       analyzed_result=AnalyzedResult(value=rng.standard_normal((L, 1)).astype(np.float32),
                                                      log_probs=rng.standard_normal((L, 1)).astype(np.float32)),
                       policy_state=NamedArray(hx=rng.standard_normal((L, 1, 16)).astype(np.float32)),
                       policy_version_steps=np.full((L, 1), 3, dtype=np.int64))


@pytest.fixture
def blosc1_writer_as_blosc(monkeypatch):
    """Encoding needs a module called `blosc`: the test writer under that name (compress only -- decoding is the library's)."""
    mod = types.ModuleType("blosc")
    mod.compress = lambda data, typesize=8, clevel=9, shuffle=1, cname="blosclz": W.compress(bytes(data), typesize, 1 << 14)

    def no_decompress(data):
        raise AssertionError("the package must not decode when the native decoder is chosen")

    mod.decompress = no_decompress
    monkeypatch.setitem(sys.modules, "blosc", mod)
    monkeypatch.setattr(wire, "_codec_choice", "native")


@pytest.mark.parametrize("method", ["raw_compress", "compress_pickle", "pickle_compress", "obs_compress",
                                    "compress_except_policy_state"])
def test_wire_round_trip_through_the_native_decoder(method, blosc1_writer_as_blosc):
    x = _sample(np.random.default_rng(5))
    x.register_metadata(birth_time=12)
    msg = wire.dumps(x, method)
    back = wire.loads(msg)
    want, got = dict(flatten(x)), dict(flatten(back))
    assert list(want) == list(got) and back.metadata == x.metadata and back.metadata["birth_time"] == 12
    for k, v in want.items():
        assert (got[k] is None) if v is None else (got[k].dtype == v.dtype and np.array_equal(got[k], v)), k


def test_lazy_frames_hand_out_compressed_leaves(blosc1_writer_as_blosc):
    x = _sample(np.random.default_rng(6))
    msg = wire.dumps(x, "obs_compress")
    entries, _ = wire.frames(msg, lazy=True)
    by_key = {k: p for k, _, _, p in entries}
    leaf = by_key["obs.frame"]
    assert isinstance(leaf, wire.CompressedLeaf) and leaf.shape == x.obs.frame.shape and leaf.dtype == np.uint8
    assert not isinstance(by_key["reward"], wire.CompressedLeaf)  # 'obs' not in the key: a view of the message
    dst = np.full(leaf.nbytes + 8, 7, dtype=np.uint8)
    leaf.decode_into(dst[4:-4], threads=2)
    assert np.array_equal(dst[4:-4].reshape(leaf.shape), x.obs.frame) and (dst[:4] == 7).all() and (dst[-4:] == 7).all()
    assert np.array_equal(np.asarray(leaf), x.obs.frame)
    with pytest.raises(ValueError, match="contiguous uint8"):
        leaf.decode_into(np.empty(leaf.nbytes - 1, dtype=np.uint8))
    # a payload whose decoded size disagrees with dtype x shape is rejected when the message is read, as for raw payloads
    keys = [k for k, *_ in entries]
    i = 1 + 4 * keys.index("obs.frame") + 2
    bad = list(msg)
    bad[i] = b"(3, 3)"
    with pytest.raises(ValueError, match="needs"):
        wire.frames(bad, lazy=True)


def test_codec_choice(monkeypatch):
    """auto: the native decoder where the package is absent; where a package is present it must reproduce the package's own
    frames first -- a package that writes something else (here: the zlib stand-in of tests/util.py) keeps the decoding."""
    import importlib.util
    from tests.util import StandInBlosc
    monkeypatch.delenv("SRL_B200_WIRE_CODEC", raising=False)
    monkeypatch.setattr(wire, "_codec_choice", None)
    if importlib.util.find_spec("blosc") is None:
        assert wire.codec() == "native"
    monkeypatch.setattr(wire, "_codec_choice", None)
    monkeypatch.setitem(sys.modules, "blosc", StandInBlosc)
    with pytest.warns(UserWarning, match="disagrees"):
        assert wire.codec() == "blosc"
    monkeypatch.setattr(wire, "_codec_choice", None)
    mod = types.ModuleType("blosc")
    mod.compress = lambda data, typesize=8, clevel=9, shuffle=1, cname="blosclz": W.compress(bytes(data), typesize)
    monkeypatch.setitem(sys.modules, "blosc", mod)
    assert wire.codec() == "native"  # a package whose frames the native decoder reproduces
    monkeypatch.setattr(wire, "_codec_choice", None)
    monkeypatch.setenv("SRL_B200_WIRE_CODEC", "blosc")
    assert wire.codec() == "blosc"
    monkeypatch.setattr(wire, "_codec_choice", None)
    monkeypatch.setenv("SRL_B200_WIRE_CODEC", "zip")
    with pytest.raises(ValueError):
        wire.codec()
