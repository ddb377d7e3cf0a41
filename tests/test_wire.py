"""CPU tests of the sample wire format (srl_b200/wire.py) against messages the UNMODIFIED reference produced
(tests/golden/wire.npz, written by oracle/make_golden.py from base.namedarray.dumps, namedarray.py:100-163)."""
import pickle

import numpy as np
import pytest

from oracle import ref_loader
from srl_b200 import wire
from srl_b200.api import AnalyzedResult, SampleBatch
from srl_b200.namedarray import NamedArray, flatten
from tests.util import StandInBlosc, load_golden

FRAMED = ["raw_bytes", "raw_compress", "compress_pickle", "obs_compress", "compress_except_policy_state"]


@pytest.fixture
def blosc_stand_in(monkeypatch):
    """`blosc` is not installed here; the compressed fixtures were written by the unmodified reference with this
    stand-in codec under that name (oracle/make_golden.py::gen_wire)."""
    import sys
    monkeypatch.setitem(sys.modules, "blosc", StandInBlosc)
    # the stand-in writes zlib streams behind a marker, not Blosc-1 frames: the package under that name decodes them, not
    # the library's native decoder (wire.codec(); tests/test_wire_native.py covers that one)
    monkeypatch.setattr(wire, "_codec_choice", "blosc")


def _message(fx, method):
    stream, lens = fx[f"{method}.stream"].tobytes(), fx[f"{method}.lens"]
    off = np.concatenate([[0], np.cumsum(lens)])
    return [stream[off[i]:off[i + 1]] for i in range(len(lens))]


def _same_sample(fx):
    """The fixture's sample rebuilt with this repo's containers."""
    leaf = lambda k: fx[f"leaf.{k}"]
    return SampleBatch(obs=NamedArray(frame=leaf("obs.frame"), vec=leaf("obs.vec")), on_reset=leaf("on_reset"),
                       done=leaf("done"), truncated=None, action=NamedArray(x=leaf("action.x")), reward=leaf("reward"),
                       analyzed_result=AnalyzedResult(value=leaf("analyzed_result.value"),
                                                      log_probs=leaf("analyzed_result.log_probs")),
                       policy_state=NamedArray(hx=leaf("policy_state.hx")),
                       policy_version_steps=leaf("policy_version_steps"), sampling_weight=2.5)


@pytest.mark.parametrize("method", FRAMED)
def test_framed_decode_of_reference_message(method, blosc_stand_in):
    fx = load_golden("wire.npz")
    x = wire.loads(_message(fx, method))
    got = dict(flatten(x))
    for k in [k[5:] for k in fx if k.startswith("leaf.")]:
        ref = fx[f"leaf.{k}"]
        want_dtype = np.uint8 if ref.dtype == np.bool_ else ref.dtype  # bool travels as uint8 (numpy_utils.py:65-66)
        assert got[k].dtype == want_dtype and got[k].shape == ref.shape
        assert np.array_equal(got[k], ref.astype(want_dtype)), k
    for k in fx["none_leaves"]:
        assert got[str(k)] is None
    assert x.metadata == dict(sampling_weight=2.5)


@pytest.mark.parametrize("method", FRAMED)
def test_framed_encode_is_byte_identical_to_reference(method, blosc_stand_in):
    """Same frames in the same order; exactly the reference's choice of compressed leaves ('obs' in key /
    'policy_state' not in key / all / none, namedarray.py:141-156) -- the stand-in codec is deterministic, so equality
    of the payload bytes means the same leaves went through it."""
    fx = load_golden("wire.npz")
    ours = wire.dumps(_same_sample(fx), method)
    ref = _message(fx, method)
    assert len(ours) == len(ref)
    if method == "compress_pickle":
        assert pickle.loads(ours[1]) == pickle.loads(ref[1])
    else:
        assert ours[:-1] == ref[:-1]
    assert pickle.loads(ours[-1]) == pickle.loads(ref[-1])
    payloads = pickle.loads(ours[1]) if method == "compress_pickle" else ours[1:-1]
    marked = {bytes(payloads[i]).decode() for i in range(0, len(payloads), 4)
              if bytes(payloads[i + 3]).startswith(StandInBlosc.MARK)}
    present = {k[5:] for k in fx if k.startswith("leaf.")}
    want = {"raw_bytes": set(), "raw_compress": present, "compress_pickle": present,
            "obs_compress": {k for k in present if "obs" in k},
            "compress_except_policy_state": {k for k in present if "policy_state" not in k}}[method]
    assert marked == want


def test_pickle_dict_decode_of_reference_message():
    fx = load_golden("wire.npz")
    x = wire.loads(_message(fx, "pickle_dict"))
    got = dict(flatten(x))
    assert np.array_equal(got["obs.frame"], fx["leaf.obs.frame"]) and got["on_reset"].dtype == np.bool_
    assert got["truncated"] is None and x.metadata == dict(sampling_weight=2.5)


@pytest.mark.parametrize("method", ["raw_bytes", "pickle_dict", "pickle", "raw_compress", "compress_pickle",
                                    "pickle_compress", "obs_compress", "compress_except_policy_state"])
def test_round_trip(method, blosc_stand_in):
    """base/tests/namedarray_test.py:261-284 (test_serialization): every encoding method round-trips data and metadata."""
    fx = load_golden("wire.npz")
    x = _same_sample(fx)
    x.register_metadata(a=1, b="xxx")
    y = wire.loads(wire.dumps(x, method))
    for (ka, va), (kb, vb) in zip(flatten(x), flatten(y)):
        assert ka == kb and ((va is None and vb is None) or np.array_equal(np.asarray(va).astype(vb.dtype), vb))
    assert y.metadata == x.metadata


def test_frames_are_zero_copy_views_and_checked():
    fx = load_golden("wire.npz")
    msg = _message(fx, "raw_bytes")
    entries, meta = wire.frames(msg)
    by_key = {k: (dt, shape, p) for k, dt, shape, p in entries}
    dt, shape, p = by_key["obs.frame"]
    i = 1 + 4 * [k for k, *_ in entries].index("obs.frame") + 3
    assert p.obj is msg[i] and p.nbytes == int(np.prod(shape)) * dt.itemsize  # a view of the received bytes
    bad = list(msg)
    bad[i] = bad[i][:-1]
    with pytest.raises(ValueError, match="payload"):
        wire.frames(bad)
    with pytest.raises(ValueError, match="multiple of 4"):
        wire.frames(msg[:3] + msg[4:])


def test_compressed_methods_need_blosc_like_the_reference(monkeypatch):
    """With the package as the decoder (SRL_B200_WIRE_CODEC=blosc) and no package installed, the compressed methods raise at
    the import, exactly where the reference's do (namedarray.py:101-103,168-171) -- encoding always does; unknown codes /
    method names raise NotImplementedError (namedarray.py:158-160,209-211)."""
    import importlib.util
    if importlib.util.find_spec("blosc") is not None:
        pytest.skip("blosc is installed here")
    monkeypatch.setattr(wire, "_codec_choice", "blosc")
    fx = load_golden("wire.npz")
    for code in (b"0004", b"0005", b"0006", b"0007", b"0008"):
        with pytest.raises(ModuleNotFoundError):
            wire.loads([code] + _message(fx, "raw_bytes")[1:])
    with pytest.raises(ModuleNotFoundError):
        wire.frames(_message(fx, "obs_compress"))
    with pytest.raises(ModuleNotFoundError):
        wire.dumps(_same_sample(fx), "obs_compress")
    with pytest.raises(NotImplementedError):
        wire.loads([b"0042", b""])
    with pytest.raises(NotImplementedError):
        wire.dumps(_same_sample(fx), "zstd")


def test_frames_of_a_compressed_message_keep_uncompressed_payloads_zero_copy(blosc_stand_in):
    fx = load_golden("wire.npz")
    msg = _message(fx, "obs_compress")
    entries, _ = wire.frames(msg)
    keys = [k for k, *_ in entries]
    i = 1 + 4 * keys.index("reward") + 3
    assert entries[keys.index("reward")][3].obj is msg[i]          # not compressed: a view of the message
    assert entries[keys.index("obs.frame")][3].obj is not msg[1 + 4 * keys.index("obs.frame") + 3]  # decompressed copy
    with pytest.raises(ValueError):  # pickle-bodied methods carry no leaf frames
        wire.frames(_message(fx, "pickle_dict"))


@pytest.mark.skipif(not ref_loader.available(), reason="live reference only in the build container")
def test_live_reference_decodes_our_raw_bytes():
    R = ref_loader.load()
    fx = load_golden("wire.npz")
    back = R.namedarray.loads(wire.dumps(_same_sample(fx), "raw_bytes"))
    got = dict(R.namedarray.flatten(back))
    assert np.array_equal(got["obs.frame"], fx["leaf.obs.frame"]) and got["truncated"] is None
    assert np.array_equal(got["analyzed_result.log_probs"], fx["leaf.analyzed_result.log_probs"])


def test_metadata_keys_must_differ_from_fields():
    """base/tests/namedarray_test.py:286-289 (test_metadata)."""
    x = NamedArray(obs=np.zeros(3), reward=np.zeros(3))
    with pytest.raises(KeyError):
        x.register_metadata(obs=3)
    x.register_metadata(a=1, b="xxx")
    assert x.metadata == dict(a=1, b="xxx") and x.pop_metadata("a") == 1 and x.metadata == dict(b="xxx")
    x.clear_metadata()
    assert x.metadata == {}
