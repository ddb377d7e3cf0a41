"""Shared helpers for the parity tests."""
import os

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLDEN = os.path.join(ROOT, "tests", "golden")

# BASELINE.json north_star: "returns, advantages, losses and gradients within 1e-5 relative in fp32",
# defined as |x - ref| <= 1e-5 * max(1, |ref|).
TOL = 1e-5


def load_golden(name):
    with np.load(os.path.join(GOLDEN, name), allow_pickle=False) as z:
        return {k: z[k] for k in z.files}


def to_np(x):
    try:
        import torch
        if isinstance(x, torch.Tensor):
            return x.detach().cpu().numpy()
    except ImportError:
        pass
    return np.asarray(x)


def assert_close_ref(x, ref, tol=TOL, what=""):
    x, ref = to_np(x).astype(np.float64), to_np(ref).astype(np.float64)
    assert x.shape == ref.shape, f"{what}: shape {x.shape} vs reference {ref.shape}"
    both_nan = np.isnan(x) & np.isnan(ref)
    err = np.abs(x - ref)
    bound = tol * np.maximum(1.0, np.abs(ref))
    bad = ~(err <= bound) & ~both_nan
    if bad.any():
        i = np.unravel_index(np.argmax(np.where(bad, err, 0)), err.shape)
        raise AssertionError(f"{what}: {bad.sum()} of {bad.size} elements outside {tol:g}*max(1,|ref|); worst at {i}: "
                             f"got {x[i]!r}, reference {ref[i]!r}")


def assert_grad_close(g, ref, mask_sum, tol=TOL, what=""):
    """Gradients of a masked MEAN are O(1/M); compare them on the O(1) scale (g * M) so the stated
    tolerance is meaningful instead of vacuous."""
    assert_close_ref(to_np(g).astype(np.float64) * mask_sum, to_np(ref).astype(np.float64) * mask_sum, tol, what)


class StandInBlosc:
    """A stand-in for the third-party `blosc` module (absent here): same call signatures as the reference uses
    (`compress(bytes, typesize=4, cname='lz4')`, `decompress(bytes)`, base/namedarray.py:126,185), zlib inside and a marker
    in front so that an uncompressed payload can never pass for a compressed one.  Injected as sys.modules['blosc'] both
    when oracle/make_golden.py lets the UNMODIFIED reference write the compressed fixtures and when the tests read them:
    what gets pinned is the framing (which leaves are compressed, order, None leaves), not the codec."""
    MARK = b"STANDIN-BLOSC:"

    @classmethod
    def compress(cls, data, typesize=8, clevel=9, shuffle=1, cname="blosclz"):
        import zlib
        assert typesize == 4 and cname == "lz4", "the reference always passes typesize=4, cname='lz4'"
        return cls.MARK + zlib.compress(bytes(data), 1)

    @classmethod
    def decompress(cls, data):
        import zlib
        data = bytes(data)
        assert data.startswith(cls.MARK), "payload was not written by compress()"
        return zlib.decompress(data[len(cls.MARK):])
