"""tcgen05 int8 limb-split contraction against exact numpy object arithmetic (bit-exact)."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _run(x, w, w_signed, lx, lw, q):
    from tools_b200 import _ffi

    x = np.ascontiguousarray(x, dtype=np.int64)
    w = np.ascontiguousarray(w, dtype=np.int64)
    out = np.empty((x.shape[0], w.shape[0]), dtype=np.int64)
    st = _ffi.lib().qf_debug_gemm_i8(_ffi.ptr(x), _ffi.ptr(w), w_signed, lx, lw, x.shape[0], w.shape[0], x.shape[1], q,
                                     _ffi.ptr(out))
    assert st == 0
    return out


@pytest.mark.parametrize("B,N,K,lx,lw,w_signed,q", [
    (5, 7, 33, 1, 1, 0, 0), (128, 128, 128, 1, 1, 1, 0), (130, 200, 1000, 2, 3, 0, 2**24),
    (300, 70, 517, 2, 3, 0, 2**24 - 3), (64, 256, 4096, 3, 2, 1, 0), (257, 300, 12352, 2, 3, 0, 2**24),
    (200, 96, 2500, 2, 4, 0, 2**32 - 5), (100, 48, 700, 5, 2, 1, 0), (33, 40, 300, 2, 8, 0, 2**61 - 1),
])
@pytest.mark.parametrize("bk", [0, 64, 128])
def test_gemm_i8_exact(B, N, K, lx, lw, w_signed, q, bk, monkeypatch):
    """bk: K block of the shared-memory operand tiles (128: SWIZZLE_128B, 64: SWIZZLE_64B, 0: the launcher's choice)."""
    if bk:
        monkeypatch.setenv("QF_I8_BLOCK_K", str(bk))
    rng = np.random.default_rng(B * N + K)
    xmax = 127 * (256**lx - 1) // 255  # largest value with lx balanced digits in [-128,127]
    x = rng.integers(-xmax, xmax + 1, (B, K), dtype=np.int64)
    if w_signed:
        wmax = 127 * (256**lw - 1) // 255
        w = rng.integers(-wmax, wmax + 1, (N, K), dtype=np.int64)
    else:
        hi = min(2 ** (8 * lw), q if q else 2 ** (8 * lw))
        w = rng.integers(0, hi, (N, K), dtype=np.int64)
    # extreme values exercise the limb boundaries
    x[0, :] = xmax
    x[-1, :] = -xmax
    want = x.astype(object).dot(w.astype(object).T)
    if q:
        want = want % q
    got = _run(x, w, w_signed, lx, lw, q)
    assert np.array_equal(got.astype(object), want)
