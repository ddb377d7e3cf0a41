"""CPU tests of the drop-in boundary: the C-ABI library builds for sm_100a, loads, and exports every
symbol include/srl_b200.h declares; the Python layer refuses to run without it (no CPU fallback)."""
import ctypes
import os
import re
import subprocess

import numpy as np
import pytest
import torch

from srl_b200 import _lib, build
from tests.util import ROOT

HEADER = os.path.join(ROOT, "include", "srl_b200.h")


@pytest.fixture(scope="module")
def lib():
    build.build()  # nvcc cross-compiles without a GPU
    return _lib.load_library()


def _declared_symbols():
    text = open(HEADER).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(srl_[a-z0-9_]+)\s*\(", text)))


def test_header_symbols_are_exported_and_bound(lib):
    names = _declared_symbols()
    assert len(names) >= 12
    for n in names:
        assert hasattr(lib, n), f"{n} is declared in include/srl_b200.h but not exported by the library"
        assert n in _lib.SIGNATURES, f"{n} has no ctypes signature in srl_b200/_lib.py"
    assert sorted(_lib.SIGNATURES) == names, "python binding lists symbols the header does not declare"


def test_library_is_sm100a_only(lib):
    out = subprocess.run(["cuobjdump", "-lelf", _lib.lib_path()], capture_output=True, text=True)
    if out.returncode != 0:
        pytest.skip("cuobjdump not available")
    archs = set(re.findall(r"sm_\d+a?", out.stdout))
    assert archs == {"sm_100a"}, archs


def test_abi_version_and_error_text(lib):
    assert lib.srl_abi_version() == _lib.ABI_VERSION
    # invalid arguments are rejected on the host, before any CUDA call: status + message, no abort
    rc = lib.srl_gae_scan(None, None, None, None, None, None, None, None, None, 1, 1, 0, 0, 0.99, 0.95, 1.0, 1.0, None,
                          None, None, None, None, None)
    assert rc == 1
    assert b"L >= 2" in lib.srl_last_error()
    rc = lib.srl_philox_perm(0, 0, 1, -1, 1, None, None)
    assert rc == 1 and b"n_env" in lib.srl_last_error()
    with pytest.raises(_lib.SrlCudaError, match="srl_batch_gather"):
        _lib.call("srl_batch_gather", None, 99, None, 1, 1, None)


def test_struct_layouts_match_header(lib):
    assert ctypes.sizeof(_lib.PpoHyper) == 7 * 8 + 4 * 4
    assert ctypes.sizeof(_lib.LeafDesc) == 48
    assert ctypes.sizeof(_lib.LossProblem) == 12 * 8
    assert lib.srl_ppo_loss_workspace_bytes(128, 4096) >= 64 + 148 * 8 * 8 * 8


def test_missing_library_fails_loudly(monkeypatch, tmp_path):
    monkeypatch.setattr(_lib, "_lib", None)
    monkeypatch.setenv("SRL_B200_LIB", str(tmp_path / "nope.so"))
    with pytest.raises(_lib.MissingCudaLibrary, match="no CPU fallback"):
        _lib.load_library()


def test_ops_refuse_cpu_tensors(lib):
    from srl_b200 import ops
    z = torch.zeros(4, 8)
    f = torch.zeros(4, 8, dtype=torch.uint8)
    with pytest.raises(ValueError, match="CUDA tensor"):
        ops.gae_scan(z, z, f, f, f, 0.99, 0.95)
    with pytest.raises(ValueError, match="CUDA tensor"):
        ops.group_stats(torch.zeros(8, 4, dtype=torch.float64))
    with pytest.raises(AssertionError, match="does not match any implemented loss"):
        ops.LossHyper(value_loss="l1").to_c()  # same failure as utils.py:246-249
    with pytest.raises(TypeError):
        ops.LossHyper(value_loss="huber", value_loss_config=dict(gamma=2)).to_c()
    hc = ops.LossHyper(value_loss="huber", value_loss_config=dict(delta=10.0), eps_clip=0.1).to_c()
    assert (hc.vl_param, hc.value_eps_clip, hc.value_loss) == (10.0, 0.1, 1)


def test_product_never_imports_the_oracle():
    """The oracle is test infrastructure: nothing under srl_b200/ may import it."""
    pkg = os.path.join(ROOT, "srl_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                src = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", src, flags=re.M), f


@pytest.mark.parametrize("nbytes", [0, 1, 4097, 300_000, (1 << 20) + 13, 9_000_001])
@pytest.mark.parametrize("threads", [1, 2, 5, 16])
def test_host_copy_is_a_memcpy_for_every_size_and_thread_count(lib, nbytes, threads):
    """srl_host_copy (host-only entry point: pure host memory, no GPU needed): byte-exact, does not touch a byte outside
    [dst, dst + bytes), unaligned pointers included."""
    rng = np.random.default_rng(nbytes + threads)
    src = rng.integers(0, 256, nbytes + 3, dtype=np.uint8)
    dst = np.full(nbytes + 64, 0xAB, dtype=np.uint8)
    _lib.call("srl_host_copy", dst.ctypes.data + 7, src.ctypes.data + 3, nbytes, threads)
    assert np.array_equal(dst[7:7 + nbytes], src[3:3 + nbytes])
    assert (dst[:7] == 0xAB).all() and (dst[7 + nbytes:] == 0xAB).all()


def test_host_copy_concurrent_callers_and_bad_arguments(lib):
    import threading
    rng = np.random.default_rng(0)
    srcs = [rng.integers(0, 256, 3_000_000 + i, dtype=np.uint8) for i in range(4)]
    dsts = [np.zeros_like(s) for s in srcs]

    def work(i):
        for _ in range(5):
            dsts[i][:] = 0
            _lib.call("srl_host_copy", dsts[i].ctypes.data, srcs[i].ctypes.data, srcs[i].nbytes, 2 + i)

    ts = [threading.Thread(target=work, args=(i,)) for i in range(4)]
    [t.start() for t in ts]
    [t.join() for t in ts]
    assert all(np.array_equal(d, s) for d, s in zip(dsts, srcs))
    with pytest.raises(_lib.SrlCudaError, match="threads"):
        _lib.call("srl_host_copy", dsts[0].ctypes.data, srcs[0].ctypes.data, 10, 0)
    with pytest.raises(_lib.SrlCudaError, match="null"):
        _lib.call("srl_host_copy", None, srcs[0].ctypes.data, 10, 1)


def test_recorded_calls_replay_the_same_entry_points(lib):
    """_lib.record_calls / bind_calls / replay_calls (HotPath.run_device(plan=True)): every call() inside the context is
    kept as (entry point, ctypes arguments) and can be issued again as it is; a failing status still raises on replay.
    Exercised with a host-only entry point (srl_host_copy), so no GPU is needed."""
    src = np.arange(4096, dtype=np.uint8) % 251
    dst = np.zeros_like(src)
    calls = []
    with _lib.record_calls(calls):
        _lib.call("srl_host_copy", dst.ctypes.data, src.ctypes.data, src.nbytes, 1)
        with pytest.raises(RuntimeError, match="does not nest"):
            with _lib.record_calls([]):
                pass
    _lib.call("srl_host_copy", dst.ctypes.data, src.ctypes.data, 16, 1)  # outside the context: not recorded
    assert [name for name, _ in calls] == ["srl_host_copy"] and np.array_equal(dst, src)
    bound = _lib.bind_calls(calls)
    src[:] = 7  # same buffers, new contents: the replay copies them again
    _lib.replay_calls(bound)
    assert (dst == 7).all()
    bad = _lib.bind_calls([("srl_host_copy", (dst.ctypes.data, src.ctypes.data, src.nbytes, 0))])  # threads = 0
    with pytest.raises(_lib.SrlCudaError, match="threads"):
        _lib.replay_calls(bad)
