"""Property tests (hypothesis) of the oracle restatement and the host-side codecs -- the size-independent properties the GPU
tests use at full size (linearity of GAE, permutation-ness of the Philox shuffle), exercised here over random shapes."""
import numpy as np
import torch
from hypothesis import given, settings, strategies as st

from oracle import ref_math as M
from srl_b200 import synth, wire
from srl_b200.namedarray import NamedArray, flatten

_FAST = settings(max_examples=25, deadline=None)


@_FAST
@given(T=st.integers(1, 40), B=st.integers(1, 9), seed=st.integers(0, 10_000), a=st.floats(-3, 3), b=st.floats(-3, 3))
def test_gae_is_linear_in_rewards_and_values(T, B, seed, a, b):
    """A_t is a linear function of (reward, value) for fixed flags (gae.py:63,87-95): the property the full-size GPU test
    checks at cfg2 (tests/test_gpu_parity.py::test_gae_full_size_property_linearity)."""
    cfg = synth.PathConfig("p", T=T, B=B, p_end=0.15)
    s = synth.make_sample_scalars(cfg, seed)
    f = {k: torch.from_numpy(s[k]).double() for k in ("done", "truncated", "on_reset")}
    g = torch.Generator().manual_seed(seed)
    alive = 1 - f["on_reset"][1:]
    r1, r2 = (torch.randn(T, B, 1, generator=g, dtype=torch.float64) * alive for _ in range(2))
    v1, v2 = (torch.randn(T + 1, B, 1, generator=g, dtype=torch.float64) for _ in range(2))
    run = lambda r, v: M.gae_trace_ref(r, v, f["truncated"], f["done"], f["on_reset"], 0.99, 0.95).double()
    lhs = run(a * r1 + b * r2, a * v1 + b * v2)
    rhs = a * run(r1, v1) + b * run(r2, v2)
    assert torch.allclose(lhs, rhs, rtol=0, atol=1e-5 * max(1.0, float(rhs.abs().max())))


@_FAST
@given(n=st.integers(1, 3000), seed=st.integers(0, 2**40), epoch=st.integers(0, 1000))
def test_philox_perm_is_a_permutation_and_depends_on_seed_and_epoch(n, seed, epoch):
    p = M.philox_perm_ref(seed, epoch, n)
    assert p.shape == (n,) and np.array_equal(np.sort(p), np.arange(n))
    if n >= 64:  # two different keys giving the same permutation of >= 64 items: probability ~ 1 / 64!
        assert not np.array_equal(p, M.philox_perm_ref(seed, epoch + 1, n))
        assert not np.array_equal(p, M.philox_perm_ref(seed + 1, epoch, n))


_DTYPES = [np.uint8, np.bool_, np.float32, np.float64, np.int32, np.int64]


@st.composite
def _records(draw, depth=0):
    fields = {}
    for name in draw(st.lists(st.sampled_from(["obs", "reward", "policy_state", "x", "y", "mask"]), min_size=1, max_size=4,
                              unique=True)):
        kind = draw(st.integers(0, 3 if depth < 2 else 2))
        if kind == 0:
            fields[name] = None
        elif kind == 3:
            fields[name] = draw(_records(depth + 1))
        else:
            shape = tuple(draw(st.lists(st.integers(0, 4), min_size=1, max_size=3)))
            dt = draw(st.sampled_from(_DTYPES))
            fields[name] = (np.arange(int(np.prod(shape))).reshape(shape) % 2).astype(dt)
    return NamedArray(**fields)


@_FAST
@given(x=_records(), method=st.sampled_from(["raw_bytes", "pickle_dict", "pickle"]))
def test_wire_round_trip_of_arbitrary_nested_records(x, method):
    """Any nesting, None leaves, empty arrays, every dtype the reference's encode_dtype accepts (numpy_utils.py:64-78)."""
    x.register_metadata(tag="t", n=3)
    y = wire.loads(wire.dumps(x, method))
    fa, fb = flatten(x), flatten(y)
    if method == "raw_bytes":
        # an all-None sub-record has no frames of its own: it comes back as leaves-with-None, same flattened keys
        assert [k for k, _ in fa] == [k for k, _ in fb]
    for (ka, va), (kb, vb) in zip(fa, fb):
        assert ka == kb
        if va is None:
            assert vb is None
        else:
            want = va.astype(np.uint8) if (va.dtype == np.bool_ and method == "raw_bytes") else va
            assert vb.dtype == want.dtype and vb.shape == want.shape and np.array_equal(vb, want)
    assert y.metadata == dict(tag="t", n=3)


def test_feistel_permutation_and_its_inverse_on_the_host(tmp_path):
    """perm.cuh on the host (its functions are __host__ __device__): perm_at is a permutation of [0, n_env) and perm_pos_of
    undoes it -- the scan kernel derives every lane's minibatch from the inverse (tests/native/perm_inverse.cu)."""
    import os
    import subprocess
    from srl_b200.build import CSRC, INCLUDE, find_nvcc
    src = os.path.join(os.path.dirname(os.path.abspath(__file__)), "native", "perm_inverse.cu")
    exe = str(tmp_path / "perm_inverse")
    subprocess.run([find_nvcc(), "-std=c++17", "-Wno-deprecated-gpu-targets", "-I", CSRC, "-I", INCLUDE, "-o", exe, src],
                   check=True, capture_output=True)
    run = subprocess.run([exe], capture_output=True, text=True)
    assert run.returncode == 0 and "bad=0" in run.stdout, run.stdout + run.stderr
