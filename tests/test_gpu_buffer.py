"""Device-resident sample buffer (srl_b200/buffer.py) against the reference's PriorityQueueBuffer semantics:
stacking bit-exact vs recursive_aggregate(np.stack(axis=1)) (base/buffer.py:118-126; golden fixture written by the
unmodified reference), LIFO / reuse / drop order as base/tests/buffer_test.py:44-73, and the trainer consuming a
device batch gives exactly what it gives for the same batch from host memory."""
import copy

import numpy as np
import pytest
import torch

from srl_b200 import api, synth
from srl_b200.namedarray import NamedArray, flatten, recursive_aggregate
from tests.util import load_golden

pytestmark = pytest.mark.gpu


def _sample(rng, L=9, with_info=True, big=False):
    obs = NamedArray(vec=rng.standard_normal((L, 5)).astype(np.float32),
                     frame=rng.integers(0, 255, (L, 4, 21, 21) if big else (L, 3, 7), dtype=np.uint8))
    return api.SampleBatch(
        obs=obs, on_reset=(rng.random((L, 1)) < 0.2).astype(np.uint8), done=(rng.random((L, 1)) < 0.1).astype(np.uint8),
        truncated=np.zeros((L, 1), dtype=np.uint8), action=NamedArray(x=rng.integers(0, 4, (L, 1)).astype(np.int32)),
        reward=rng.standard_normal((L, 1)).astype(np.float32),
        info=NamedArray(ret=rng.standard_normal((L, 1)).astype(np.float32)) if with_info else None,
        info_mask=(rng.random((L, 1)) < 0.3).astype(np.float32),
        policy_version_steps=np.full((L, 1), 7, dtype=np.int64),
        analyzed_result=api.AnalyzedResult(value=rng.standard_normal((L, 1)).astype(np.float32),
                                           log_probs=-rng.random((L, 1)).astype(np.float32)))


def test_stack_is_bit_exact_vs_np_stack():
    from srl_b200.buffer import DeviceSlabBuffer
    rng = np.random.default_rng(0)
    B = 6
    samples = [_sample(rng, big=True) for _ in range(B)]
    buf = DeviceSlabBuffer(max_size=4, reuses=1, batch_size=B)
    formed = [buf.put(copy.deepcopy(s)) for s in samples]
    assert formed == [False] * (B - 1) + [True] and buf.qsize() == 1
    entry = buf.get()
    assert buf.empty() and entry.reuses == 1 and entry.reuses_left == 0
    want = recursive_aggregate(samples, lambda xs: np.stack(xs, axis=1))
    got = dict(flatten(entry.sample))
    for k, v in flatten(want):
        if k == "trainer_worker_recv_timestamp":
            assert got[k].shape == (9, B, 1) and got[k].dtype == torch.int64 and int(got[k].min()) > 1_600_000_000
            continue
        if v is None:
            assert got[k] is None, k
            continue
        assert got[k].is_cuda and tuple(got[k].shape) == v.shape, k
        assert np.array_equal(got[k].cpu().numpy(), v), f"{k}: device stack differs from np.stack(axis=1)"
        assert got[k].cpu().numpy().dtype == v.dtype, k


def test_stack_matches_reference_fixture():
    """The same fixture the oracle is pinned with (tests/golden/stack.npz, written by the unmodified reference's
    recursive_aggregate): per-sample leaves in (`truncated` is None in every other sample), stacked leaves out."""
    from srl_b200.buffer import DeviceSlabBuffer
    from srl_b200.namedarray import from_flattened
    d = load_golden("stack.npz")
    B = len({k.split(".")[0] for k in d if k.startswith("s")})
    leaves = sorted({k.split(".", 1)[1] for k in d if k.startswith("out.")})
    buf = DeviceSlabBuffer(batch_size=B)
    for j in range(B):
        entries = [(n, d[f"s{j}.{n}"] if f"s{j}.{n}" in d else None) for n in leaves]
        buf.put(from_flattened(entries + [("trainer_worker_recv_timestamp", None)]))
    out = dict(flatten(buf.get().sample))
    for n in leaves:
        want = d[f"out.{n}"]
        got = out[n].cpu().numpy()
        assert got.dtype == want.dtype and np.array_equal(got, want), n


def test_lifo_reuse_and_drop_semantics():
    """base/tests/buffer_test.py:44-73 on device batches."""
    from srl_b200.buffer import DeviceSlabBuffer

    def one(i):
        return NamedArray(on_reset=np.zeros((3, 1), dtype=np.uint8), x=np.full((3, 1), i, dtype=np.float32),
                          trainer_worker_recv_timestamp=None)

    b = DeviceSlabBuffer(max_size=5, reuses=2, batch_size=1)
    assert b.empty()
    b.put(one(42))
    assert b.qsize() == 1
    assert b.get().sample.x[:, 0, 0].tolist() == [42, 42, 42]
    assert b.qsize() == 1  # put back: one reuse left
    assert b.get().sample.x[:, 0, 0].tolist() == [42, 42, 42]
    assert b.qsize() == 0 and b.empty()
    for i in range(10):  # dropping: only the newest five survive
        b.put(one(i))
    assert b.qsize() == 5
    for _ in range(2):
        for i in range(9, 4, -1):  # LIFO, each served twice
            assert float(b.get().sample.x[0, 0, 0]) == i
    with pytest.raises(AssertionError):
        b.get()
    nb = DeviceSlabBuffer(max_size=5, reuses=1, batch_size=0)  # no batching: objects pass through
    nb.put("Some things cannot be batched")
    assert nb.get().sample == "Some things cannot be batched"


def test_missing_leaf_is_zero_filled():
    """A leaf that is None in some samples is zero-filled there (base/namedarray.py:588-595)."""
    from srl_b200.buffer import DeviceSlabBuffer
    rng = np.random.default_rng(3)
    a, b_, c = _sample(rng), _sample(rng), _sample(rng)
    b_.info_mask = None
    buf = DeviceSlabBuffer(batch_size=3)
    for s in (a, b_, c):
        buf.put(s)
    out = buf.get().sample
    m = out.info_mask.cpu().numpy()
    assert np.array_equal(m[:, 0], a.info_mask) and np.array_equal(m[:, 2], c.info_mask) and not m[:, 1].any()


def test_trainer_consumes_device_batch():
    """MultiAgentPPOB200.step on a DeviceSlabBuffer batch == the same step on the host-stacked batch."""
    from srl_b200.buffer import DeviceSlabBuffer
    from tests.doubles import TinyActorCriticPolicy
    from srl_b200.trainer import MultiAgentPPOB200
    from tests.test_gpu_trainer import NUM_ACTIONS, OBS_DIM, make_sample
    cfg = synth.PathConfig("buf", T=12, B=8, p_end=0.08)
    kw = dict(clip_value=True, dual_clip=False, value_loss="huber", value_loss_config=dict(delta=10.0), ppo_epochs=2,
              num_minibatches=2, optimizer="sgd", optimizer_config=dict(lr=0.05), recompute_adv_on_reuse=False)
    host_batches = [make_sample(cfg, seed=20 + it) for it in range(2)]
    pol_a = TinyActorCriticPolicy(OBS_DIM, NUM_ACTIONS, device="cuda:0", seed=3)
    pol_b = TinyActorCriticPolicy(OBS_DIM, NUM_ACTIONS, device="cuda:0", seed=3)
    tr_host = MultiAgentPPOB200(pol_a, prefetch=False, **kw)
    tr_dev = MultiAgentPPOB200(pol_b, prefetch=False, **kw)
    buf = DeviceSlabBuffer(max_size=4, reuses=2, batch_size=cfg.B)
    for hb in host_batches:
        for j in range(cfg.B):  # the actor workers' per-environment samples: column j of the batch
            buf.put(hb[:, j])
        entry = buf.get()
        got = tr_dev.step(entry.sample)
        want = tr_host.step(copy.deepcopy(hb))
        for k, v in want.stats.items():
            assert got.stats[k] == pytest.approx(v, rel=1e-6, abs=1e-7), k
        assert entry.sample.analyzed_result.adv.is_cuda  # cached on the device for the next reuse
        # the entry is served again (reuses=2) with its cached advantages
        again = buf.get()
        assert again.sample is entry.sample and again.reuses == 2
        tr_dev.step(again.sample)
        tr_host.step(_with_cached(hb, entry.sample))
    for (n1, p1), (n2, p2) in zip(pol_a.net.state_dict().items(), pol_b.net.state_dict().items()):
        assert torch.allclose(p1, p2, rtol=1e-6, atol=1e-7), n1


def _with_cached(host_batch, dev_batch):
    hb = copy.deepcopy(host_batch)
    hb.analyzed_result.adv = dev_batch.analyzed_result.adv.cpu().numpy()
    hb.analyzed_result.ret = dev_batch.analyzed_result.ret.cpu().numpy()
    return hb


def test_put_frames_equals_put_of_decoded_sample():
    """Wire decode straight into the staging block: a batch assembled from raw_bytes messages
    (base/namedarray.py:115-128) is bit-identical to the batch assembled from the decoded samples, including a leaf that
    is None in some messages (zero-filled, namedarray.py:588-595) and the metadata of the first sample."""
    from srl_b200 import wire
    from srl_b200.buffer import DeviceSlabBuffer
    rng = np.random.default_rng(5)
    B = 5
    samples = [_sample(rng, big=True, with_info=(j % 2 == 0)) for j in range(B)]
    for s in samples:
        s.register_metadata(sampling_weight=1.5)
    a, b = DeviceSlabBuffer(batch_size=B), DeviceSlabBuffer(batch_size=B)
    formed = [b.put_frames(wire.dumps(s, "raw_bytes")) for s in samples]
    for s in samples:
        a.put(copy.deepcopy(s))
    assert formed == [False] * (B - 1) + [True]
    xa, xb = a.get().sample, b.get().sample
    fa, fb = dict(flatten(xa)), dict(flatten(xb))
    assert sorted(fa) == sorted(fb)
    for k, v in fa.items():
        if v is None:
            assert fb[k] is None, k
        elif k == "trainer_worker_recv_timestamp":
            assert fb[k].shape == v.shape and fb[k].dtype == v.dtype
        else:
            assert fb[k].dtype == v.dtype and torch.equal(fb[k], v), k
    assert not fb["info.ret"][:, 1].any() and fb["info.ret"][:, 0].any()  # zero-filled where the message had None
    assert xb.metadata == dict(sampling_weight=1.5)
    with pytest.raises(ValueError):  # only raw_bytes messages carry leaf frames
        b.put_frames(wire.dumps(samples[0], "pickle_dict"))


@pytest.mark.parametrize("method,copy_threads", [("obs_compress", 1), ("raw_compress", 4), ("compress_except_policy_state", 1),
                                                 ("compress_pickle", 2)])
def test_put_frames_decodes_compressed_leaves_into_the_pinned_block(method, copy_threads, monkeypatch):
    """a compressed message's payloads are decoded by the library's own Blosc-1 / LZ4 decoder
    (srl_blosc1_decompress) straight into the pinned staging block -- the batch equals the batch assembled from the samples
    themselves, bit for bit, including a string leaf (compressed on the wire, kept on the host) and a leaf that is None in
    some messages.  The frames are written by tests/blosc1_writer.py under the name `blosc` (no blosc here: see
    tests/test_wire_native.py for what that pins); the package's decompress must never be called."""
    import sys
    import types
    from srl_b200 import wire
    from srl_b200.buffer import DeviceSlabBuffer
    from tests import blosc1_writer as W
    mod = types.ModuleType("blosc")
    mod.compress = lambda data, typesize=8, clevel=9, shuffle=1, cname="blosclz": W.compress(bytes(data), typesize, 1 << 16)

    def no_decompress(data):
        raise AssertionError("the package must not decode")

    mod.decompress = no_decompress
    monkeypatch.setitem(sys.modules, "blosc", mod)
    monkeypatch.setattr(wire, "_codec_choice", "native")
    rng = np.random.default_rng(11)
    B = 4
    samples = [_sample(rng, big=True, with_info=(j % 2 == 0)) for j in range(B)]
    for j, s in enumerate(samples):
        s.policy_name = np.full((s.on_reset.shape[0], 1), f"policy_{j % 2}")
    a, b = DeviceSlabBuffer(batch_size=B), DeviceSlabBuffer(batch_size=B, copy_threads=copy_threads)
    assert [b.put_frames(wire.dumps(s, method)) for s in samples] == [False] * (B - 1) + [True]
    for s in samples:
        a.put(copy.deepcopy(s))
    fa, fb = dict(flatten(a.get().sample)), dict(flatten(b.get().sample))
    assert sorted(fa) == sorted(fb)
    for k, v in fa.items():
        if v is None:
            assert fb[k] is None, k
        elif k == "trainer_worker_recv_timestamp":
            assert fb[k].shape == v.shape and fb[k].dtype == v.dtype
        elif isinstance(v, np.ndarray):
            assert isinstance(fb[k], np.ndarray) and np.array_equal(fb[k], v), k
        else:
            assert fb[k].dtype == v.dtype and torch.equal(fb[k], v), k


def test_string_and_empty_leaves_stay_on_the_host():
    """Real SRL samples carry `policy_name` as a '<U..' string array (policy_worker.py:186; the wire format ships it,
    base/namedarray.py:115-128) and the reference's buffer stacks it like any other leaf before the trainer worker clears
    it (trainer_worker.py:169).  Such leaves -- and zero-size ones -- cannot live in HBM: they are stacked on the host with
    np.stack(axis=1) while the numeric leaves go through the device gather."""
    from srl_b200.buffer import DeviceSlabBuffer
    rng = np.random.default_rng(3)
    B, L = 4, 9
    samples = []
    for i in range(B):
        s = _sample(rng, L=L)
        s.policy_name = np.full((L, 1), f"policy_{i % 2}")
        s.send_timestamp = np.zeros((L, 0), dtype=np.int64)  # a zero-size leaf
        samples.append(s)
    buf = DeviceSlabBuffer(max_size=4, reuses=1, batch_size=B)
    assert [buf.put(copy.deepcopy(s)) for s in samples] == [False] * (B - 1) + [True]
    got = buf.get().sample
    assert isinstance(got.policy_name, np.ndarray) and got.policy_name.shape == (L, B, 1)
    assert np.array_equal(got.policy_name, np.stack([s.policy_name for s in samples], axis=1))
    assert isinstance(got.send_timestamp, np.ndarray) and got.send_timestamp.shape == (L, B, 0)
    assert got.reward.is_cuda and np.array_equal(got.reward.cpu().numpy(), np.stack([s.reward for s in samples], axis=1))
